// k-nearest-neighbour graph over C-alpha coordinates.
// Replaces CA_ProteinFeatures._dist (reference models/protein_mpnn_utils.py:447-459) without
// ever materialising the [L, L] distance matrix.
//
// One warp per query row i.  The frame's coordinates are staged once per CTA in shared memory;
// each lane forms the distances to j = lane, lane+32, ... with the oracle's exact operation
// order -- ((dx*dx + dy*dy) + dz*dz) + 1e-6, correctly rounded sqrt, no FMA contraction --
// as a 64-bit key (distance bits << 32 | j).  Selection is exact and in two steps:
//   1. pruning: the K residues nearest in SEQUENCE (a window around i) are K valid candidates, so
//      the largest of their distances T bounds the K-th smallest distance from above; only keys
//      with distance <= T are kept (compacted in order into the warp's shared-memory row --
//      typically 1-3 K of the L keys for a folded chain);
//   2. K rounds of "smallest key greater than the previous one" over the kept keys (two
//      redux.sync minima per round: distance bits, then index among the lanes that hold that
//      distance), which yields the neighbours sorted ascending with the lowest-index tie-break.
// Rows with padding in play (fewer than K valid residues, or a masked query row) keep every key:
// the reference's D + (1 - mask) * D_max puts the masked columns at the row maximum.
#include "model.h"

namespace cb2 {

__global__ void knn_topk_kernel(const float* __restrict__ X, const int* __restrict__ lengths, int L, int K,
                                float* __restrict__ D_out, int* __restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5;
    float* sX = reinterpret_cast<float*>(smem_raw);                       // [L*3]
    unsigned long long* sKeys = reinterpret_cast<unsigned long long*>(smem_raw + (((size_t)L * 12 + 15) & ~(size_t)15));
    const int f = blockIdx.y;
    const int n = lengths ? lengths[f] : L;
    const float* Xf = X + (size_t)f * L * 3;
    for (int t = threadIdx.x; t < L * 3; t += blockDim.x) sX[t] = Xf[t];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long* keys = sKeys + (size_t)warp * L;
    for (int i = blockIdx.x * warps + warp; i < L; i += gridDim.x * warps) {
        const float xi = sX[i * 3 + 0], yi = sX[i * 3 + 1], zi = sX[i * 3 + 2];
        const bool vi = i < n;
        auto dist = [&](int j) {
            float dx = __fsub_rn(sX[j * 3 + 0], xi), dy = __fsub_rn(sX[j * 3 + 1], yi), dz = __fsub_rn(sX[j * 3 + 2], zi);
            float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            return __fsqrt_rn(__fadd_rn(s, 1e-6f));
        };
        const bool prune = vi && n >= K && K < L;
        int count = L;                                     // keys kept in the warp's row
        if (prune) {
            // K consecutive valid residues around i
            const int w0 = min(max(i - K / 2, 0), n - K);
            float t = 0.0f;
            for (int j = w0 + lane; j < w0 + K; j += 32) t = fmaxf(t, dist(j));
            t = warp_max(t);
            count = 0;
            for (int base = 0; base < n; base += 32) {
                const int j = base + lane;
                float d = 0.0f;
                bool keep = false;
                if (j < n) { d = dist(j); keep = d <= t; }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep) keys[count + __popc(m & ((1u << lane) - 1u))] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
                count += __popc(m);
            }
        } else {
            float dmax = 0.0f;
            for (int j = lane; j < L; j += 32) {
                float d = 0.0f;
                if (vi && j < n) d = dist(j);
                dmax = fmaxf(dmax, d);
                keys[j] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
            }
            dmax = warp_max(dmax);
            if (n < L) {   // D_adjust = D + (1 - mask2D) * D_max
                for (int j = lane; j < L; j += 32)
                    if (!(vi && j < n)) keys[j] = ((unsigned long long)__float_as_uint(dmax) << 32) | (unsigned)j;
            }
        }
        __syncwarp();
        unsigned long long prev = 0ull;
        bool first = true;
        float* Drow = D_out + ((size_t)f * L + i) * K;
        int* Irow = idx_out + ((size_t)f * L + i) * K;
        for (int k = 0; k < K; ++k) {
            unsigned long long best = ~0ull;
            for (int j = lane; j < count; j += 32) {
                unsigned long long key = keys[j];
                if ((first || key > prev) && key < best) best = key;
            }
            // warp minimum of the 64-bit keys: distance bits first, then the index among the lanes that hold that distance
            const unsigned hi = (unsigned)(best >> 32);
            const unsigned hmin = __reduce_min_sync(0xffffffffu, hi);
            const unsigned lmin = __reduce_min_sync(0xffffffffu, hi == hmin ? (unsigned)(best & 0xffffffffu) : 0xffffffffu);
            best = ((unsigned long long)hmin << 32) | lmin;
            if (lane == 0) {
                Drow[k] = __uint_as_float(hmin);
                Irow[k] = (int)lmin;
            }
            prev = best;
            first = false;
        }
        __syncwarp();
    }
}

int launch_knn(const float* X, const int* lengths, int F, int L, int K, float* D, int* idx, cudaStream_t s) {
    if (K > L) { set_error("knn: K=%d > L=%d", K, L); return (int)cudaErrorInvalidValue; }
    int warps = 8;
    auto need = [&](int w) { return (((size_t)L * 12 + 15) & ~(size_t)15) + (size_t)w * L * 8; };
    while (warps > 1 && need(warps) > 200 * 1024) warps >>= 1;
    size_t smem = need(warps);
    if (smem > 227 * 1024) { set_error("knn: L=%d does not fit shared memory", L); return (int)cudaErrorInvalidValue; }
    CB2_CUDA(cudaFuncSetAttribute(knn_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int gx = (L + warps - 1) / warps;
    if (gx > 4096) gx = 4096;
    dim3 grid(gx, F);
    knn_topk_kernel<<<grid, warps * 32, smem, s>>>(X, lengths, L, K, D, idx);
    CB2_LAUNCH_CHECK();
    return 0;
}

}  // namespace cb2
