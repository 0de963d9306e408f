"""Clean-room stand-ins for the third-party packages the reference imports and this image
lacks.  TEST INFRASTRUCTURE; parity unpinned (see oracle/__init__.py)."""
