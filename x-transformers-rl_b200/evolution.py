"""LatentGenePool with the reference's interface (x_transformers_rl/evolution.py:28-184); `evolve_` is
ONE launch of xtrl_evolve (csrc/evolve.cu).  Host code only draws the three noise tensors with torch,
in the reference's order and shapes (evolution.py:113,133,173), and syncs the seed across ranks."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn, tensor

from . import ops
from .distributed import is_distributed, sync_seed


class LatentGenePool(nn.Module):
    def __init__(self, dim, num_genes_per_island, num_selected, tournament_size, num_elites=1,
                 mutation_std_dev=0.1, num_islands=1, migrate_genes_every=10, num_frac_migrate=0.1):
        super().__init__()
        assert num_islands >= 1
        assert num_genes_per_island > 2
        assert 2 <= num_selected < num_genes_per_island, 'must select at least 2 genes for mating'
        assert 0. <= num_frac_migrate <= 1.
        self.num_islands, self.num_genes_per_island = num_islands, num_genes_per_island
        self.num_genes = num_genes_per_island * num_islands
        self.num_selected, self.num_children = num_selected, num_genes_per_island - num_selected
        self.tournament_size, self.dim_gene = tournament_size, dim
        self.num_elites, self.mutation_std_dev = num_elites, mutation_std_dev
        self.num_frac_migrate, self.migrate_genes_every = num_frac_migrate, migrate_genes_every
        self.genes = nn.Parameter(F.normalize(torch.randn(self.num_genes, dim), dim=-1), requires_grad=False)
        self.register_buffer('step', tensor(0))
        self._step_host = 0          # mirrors `step` so evolve_ needs no device->host sync

    def __getitem__(self, idx):
        return F.normalize(self.genes[idx], dim=-1)

    def draw_noise(self, device):
        """the reference's randn calls, in its order (evolution.py:113, :133, :173)."""
        I, c, s, G, g = self.num_islands, self.num_children, self.num_selected, self.dim_gene, self.num_genes_per_island
        nt = torch.randn((I, c, s), device=device)
        nc = torch.randn((I, c, G), device=device)
        nm = torch.randn((I, g - 1 if self.num_elites > 0 else g, G), device=device) if self.mutation_std_dev > 0 else None
        return nt, nc, nm

    @torch.no_grad()
    def evolve_(self, fitnesses, temperature=1.5, noise=None):
        device = fitnesses.device
        assert fitnesses.ndim == 1 and fitnesses.shape[0] == self.num_genes
        if is_distributed():
            torch.manual_seed(sync_seed(device))
        nt, nc, nm = self.draw_noise(device) if noise is None else noise
        migrate = ((self._step_host + 1) % self.migrate_genes_every == 0 and self.num_islands > 1
                   and self.num_frac_migrate > 0.)
        num_migrate = 0
        if migrate:
            n_can = self.num_selected - (1 if self.num_elites > 0 else 0)
            num_migrate = max(1, n_can * self.num_frac_migrate)
            if not isinstance(num_migrate, int):      # reference evolution.py:150-154 slices with a float here
                raise TypeError('slice indices must be integers or None or have an __index__ method '
                                '(reference behaviour when num_can_migrate * num_frac_migrate > 1)')
        sel = ops.evolve_(self.genes.data, fitnesses.float().contiguous(), nt, nc, nm, num_islands=self.num_islands,
                          num_genes_per_island=self.num_genes_per_island, num_selected=self.num_selected,
                          tournament_size=self.tournament_size, num_elites=self.num_elites,
                          mutation_std_dev=self.mutation_std_dev, migrate=migrate, num_migrate=num_migrate,
                          temperature=temperature)
        self.step.add_(1)
        self._step_host += 1
        return sel
