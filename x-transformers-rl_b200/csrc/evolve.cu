// evolve.cu -- one evolution step of the latent gene pool, one CTA per island.
//
// Replaces LatentGenePool.evolve_ (x_transformers_rl/evolution.py:76-184): per-island fitness sort
// (:101), selection (:103-107), tournaments over the fitness-sorted selection (:113-127, including
// the slot-index-as-gene-index behaviour), lerp crossover (:131-133), optional
// migration (:137-160), mutation of all but the elite (:168-176) and re-normalisation (:178-180).
// The reference's three randn draws are inputs so selection is reproducible.
//
// Integer work (ranks by counting) in shared memory; gene rows move as coalesced warp rows.  In place: the s selected rows
// of the island -- the only genes any output row depends on -- are staged in shared memory first, then every row of the
// island is rebuilt, mutated, normalised and written straight over the pool (each gene byte read at most once and written
// once; no scratch round trip).  HBM-bound in principle (2*N*G*4 B + noise), latency-bound at 64 genes -- one launch, no
// host sync.
#include "common.cuh"

namespace xtrl {


__device__ __forceinline__ float torch_lerp(float a, float b, float w) {
    const float d = b - a;                       // ATen Lerp.h
    return w < 0.5f ? fmaf(w, d, a) : fmaf(-d, 1.f - w, b);
}

// NPL: gene elements per lane kept in registers between the sum of squares and the normalised write (G <= 32 * NPL);
// 0 = any G, the row is formed twice instead.  The block size is chosen by the launcher (small islands: small blocks, so
// that thousands of islands are resident at once; one big island: up to 1024 threads).
template <int NPL>
__global__ void __launch_bounds__(1024)
evolve_kernel(float *__restrict__ genes, const float *__restrict__ fitness,
              const float *__restrict__ noise_tourn, const float *__restrict__ noise_cross,
              const float *__restrict__ noise_mut, int64_t *__restrict__ selected_ids,
              float *__restrict__ scratch, const xtrl_pool_desc pd) {
    extern __shared__ unsigned char smem_raw[];
    const int g = pd.num_genes_per_island, s = pd.num_selected, k = pd.tournament_size, G = pd.dim;
    const int c = g - s;
    float *fit = (float *)smem_raw;              // [g]
    float *sfit = fit + g;                       // [g] descending
    int *order = (int *)(sfit + g);              // [g] gene index at each rank
    int *par1 = order + g;                       // [c] parent slots (evolution.py:120)
    int *par2 = par1 + c;                        // [c]
    int *src = par2 + c;                         // [s] selected row feeding output row j
    float *sel = (float *)(src + s + ((g + 2 * c + s) & 1));      // [s][G] the selected rows in fitness order (8-byte aligned)

    const int isl = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *f_in = fitness + (int64_t)isl * g;
    float *gene_rows = genes + (int64_t)isl * g * G;
    (void)scratch;

    for (int i = tid; i < g; i += (int)blockDim.x) fit[i] = f_in[i];
    __syncthreads();
    // stable descending ranks by counting (:101)
    for (int i = tid; i < g; i += (int)blockDim.x) {
        const float fi = fit[i];
        int r = 0;
        for (int j = 0; j < g; ++j) {
            const float fj = fit[j];
            r += (fj > fi) || (fj == fi && j < i);
        }
        order[r] = i;
        sfit[r] = fi;
    }
    __syncthreads();
    for (int j = tid; j < s; j += (int)blockDim.x) selected_ids[(int64_t)isl * s + j] = order[j];

    // tournaments: the k smallest noise values of each child's row pick k ranks of the selection;
    // the two fittest of them (by fitness of that rank) give the parent *slots* (:113-120)
    for (int ch = tid; ch < c; ch += (int)blockDim.x) {
        const float *nz = noise_tourn + ((int64_t)isl * c + ch) * s;
        float best1 = -INFINITY, best2 = -INFINITY;
        int slot1 = 0, slot2 = 0;
        for (int a = 0; a < s; ++a) {
            const float na = nz[a];
            int r = 0;
            for (int b2 = 0; b2 < s; ++b2) {
                const float nb = nz[b2];
                r += (nb < na) || (nb == na && b2 < a);
            }
            if (r < k) {                         // tourn_ids[ch][r] = a; its fitness is sfit[a]
                const float tf = sfit[a];
                if (tf > best1) { best2 = best1; slot2 = slot1; best1 = tf; slot1 = r; }
                else if (tf > best2) { best2 = tf; slot2 = r; }
            }
        }
        par1[ch] = slot1;
        par2[ch] = slot2;
    }
    // migration permutation of the selected rows (:137-160)
    const int has_elites = pd.num_elites > 0;
    for (int j = tid; j < s; j += (int)blockDim.x) {
        int sj = j;
        if (pd.migrate) {
            const int base = has_elites ? 1 : 0, n = s - base, nm = pd.num_migrate;
            if (j >= base) {
                const int q = j - base;          // position inside the migratable part
                if (q < nm) sj = base + (n - nm) + q;                    // tail first
                else { const int h = n - nm, p = q - nm; sj = base + (p - 1 + h) % h; }   // roll by 1 (h > 0 here)
            }
        }
        src[j] = sj;
    }
    __syncthreads();

    // stage the selected rows (rank order) -- after this nothing reads the pool any more
    for (int i = tid; i < s * G; i += (int)blockDim.x) {
        const int j = i / G, d = i - j * G;
        sel[i] = gene_rows[(int64_t)order[j] * G + d];
    }
    __syncthreads();

    // build, mutate and normalise every output row in place: warp per row, lanes across the gene dim; the row is formed
    // twice (sum of squares, then the normalised write) so that no copy of it has to live anywhere
    const float std_dev = pd.mutation_std_dev;
    for (int row = warp; row < g; row += (int)(blockDim.x >> 5)) {
        auto value = [&](int d) {
            float x;
            if (row < s) {
                x = sel[src[row] * G + d];
            } else {
                const int ch = row - s;
                const float a = sel[par1[ch] * G + d], b = sel[par2[ch] * G + d];
                const float z = noise_cross[((int64_t)isl * c + ch) * G + d] / pd.temperature;
                x = torch_lerp(a, b, 1.f / (1.f + expf(-z)));
            }
            if (std_dev > 0.f) {
                if (!has_elites) x += noise_mut[((int64_t)isl * g + row) * G + d] * std_dev;
                else if (row > 0) x += noise_mut[((int64_t)isl * (g - 1) + row - 1) * G + d] * std_dev;
            }
            return x;
        };
        float ss = 0.f;
        if (NPL > 0) {
            float xs[NPL > 0 ? NPL : 1];
#pragma unroll
            for (int i = 0; i < NPL; ++i) { const int d = lane + 32 * i; xs[i] = d < G ? value(d) : 0.f; ss = fmaf(xs[i], xs[i], ss); }
            ss = warp_sum(ss);
            const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);     // F.normalize eps
#pragma unroll
            for (int i = 0; i < NPL; ++i) { const int d = lane + 32 * i; if (d < G) gene_rows[(int64_t)row * G + d] = xs[i] * inv; }
        } else {
            for (int d = lane; d < G; d += 32) { const float x = value(d); ss = fmaf(x, x, ss); }
            ss = warp_sum(ss);
            const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            for (int d = lane; d < G; d += 32) gene_rows[(int64_t)row * G + d] = value(d) * inv;
        }
    }
}

}  // namespace xtrl

extern "C" int xtrl_evolve(float *genes, const float *fitness, const float *noise_tourn,
                           const float *noise_cross, const float *noise_mut, int64_t *selected_ids,
                           float *scratch, const xtrl_pool_desc *pd, void *stream) {
    using namespace xtrl;
    XTRL_REQUIRE(pd, XTRL_E_BADARG, "xtrl_evolve: null desc");
    XTRL_REQUIRE(genes && fitness && noise_tourn && noise_cross && selected_ids, XTRL_E_BADARG,
                 "xtrl_evolve: null pointer");
    XTRL_REQUIRE(pd->num_islands >= 1 && pd->num_genes_per_island > 2 && pd->dim >= 1, XTRL_E_BADARG,
                 "xtrl_evolve: bad pool shape");
    XTRL_REQUIRE(pd->num_selected >= 2 && pd->num_selected < pd->num_genes_per_island, XTRL_E_BADARG,
                 "xtrl_evolve: must select at least 2 and fewer than all genes (evolution.py:52)");
    XTRL_REQUIRE(pd->tournament_size >= 2 && pd->tournament_size <= pd->num_selected, XTRL_E_BADARG,
                 "xtrl_evolve: tournament_size=%d must be in [2, num_selected=%d]", pd->tournament_size, pd->num_selected);
    XTRL_REQUIRE(pd->mutation_std_dev <= 0.f || noise_mut, XTRL_E_BADARG, "xtrl_evolve: mutation needs noise_mut");
    if (pd->migrate) {
        const int n = pd->num_selected - (pd->num_elites > 0 ? 1 : 0);
        XTRL_REQUIRE(pd->num_migrate >= 1 && pd->num_migrate <= n, XTRL_E_BADARG,
                     "xtrl_evolve: num_migrate=%d must be in [1, %d]", pd->num_migrate, n);
    }
    const int g = pd->num_genes_per_island, s = pd->num_selected;
    const size_t smem = sizeof(float) * 2 * g + sizeof(int) * (g + 2 * (g - s) + s + 1) + sizeof(float) * (size_t)s * pd->dim;
    XTRL_REQUIRE(smem <= 200 * 1024, XTRL_E_UNSUPPORTED, "xtrl_evolve: an island of %d genes with %d selected rows of %d floats needs %zu B shared memory (max 200 KiB)",
                 g, s, pd->dim, smem);
    // block size: many small islands -> small blocks (all islands resident at once); few islands -> many warps per island
    int threads = 32 * ((g + 7) / 8);
    if (pd->num_islands < 2 * num_sms()) threads = 32 * g;
    threads = threads < 64 ? 64 : (threads > 1024 ? 1024 : threads);
    const int npl = pd->dim <= 128 ? 4 : (pd->dim <= 256 ? 8 : 0);
    auto kern = npl == 4 ? evolve_kernel<4> : (npl == 8 ? evolve_kernel<8> : evolve_kernel<0>);
    if (smem > 48 * 1024) XTRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<pd->num_islands, threads, smem, (cudaStream_t)stream>>>(
        genes, fitness, noise_tourn, noise_cross, noise_mut, selected_ids, scratch, *pd);
    XTRL_LAUNCH_CHECK("evolve_kernel");
    return XTRL_OK;
}
