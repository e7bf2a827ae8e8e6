// Edge featuriser: CA_ProteinFeatures.forward (reference models/protein_mpnn_utils.py:478-523)
// followed by W_e (models/latent_model.py:216), computed per (i, k) neighbour pair straight from
// the three C-alpha positions around each endpoint -- the reference builds nine dense [L, L]
// distance matrices and gathers from them.  Depends only on the C-alpha trace, so the plan runs
// it once per frame and every ensemble member / diffusion step reuses h_E0.
//
// One CTA per node i (its K <= 64 neighbour rows), 256 threads:
//   phase A  raw[r] = [16 Gaussians x 9 distances | 3 direction + 4 quaternion]   (shared memory)
//   phase B  E_pre = raw @ W_edge[:, 16:]^T + PT[clip(i - j + 32, 0, 64)]           (PT folds the
//            one-hot(66) -> Linear(66,16) positional embedding through W_edge[:, :16])
//   phase C  E = LayerNorm_affine(E_pre, eps 1e-5);  h_E0 = E @ W_e^T + b_e
#include "model.h"
#include "tile_gemm.cuh"

namespace cb2 {

namespace {

constexpr int RAW_LD = 152;   // 144 + 7 + 1 pad (multiple of 4 for float4 reads)
constexpr int MAXK = 64;

__constant__ float c_rbf_mu[16];

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 unit(V3 a) {   // F.normalize: v / max(|v|, 1e-12)
    float n = fmaxf(sqrtf(dot(a, a)), 1e-12f);
    return {a.x / n, a.y / n, a.z / n};
}
__device__ __forceinline__ V3 load3(const float* X, int i, int L) {
    if (i < 0 || i >= L) return {0.f, 0.f, 0.f};
    return {X[i * 3 + 0], X[i * 3 + 1], X[i * 3 + 2]};
}
__device__ __forceinline__ V3 gated_unit_bond(V3 a, V3 b) {   // unit(b - a) if 3.6 < |b - a| < 4.0 else 0
    V3 d = sub(b, a);
    float n = sqrtf(dot(d, d));
    if (!(3.6f < n && n < 4.0f)) return {0.f, 0.f, 0.f};
    return unit(d);
}

// Local frame rows (o1, n2, o1 x n2); defined for 1 <= i <= L-3 of the PADDED length
// (protein_mpnn_utils.py:398-426, F.pad(O, (0,0,1,2))).
__device__ void residue_frame(const float* X, int i, int L, float O[9]) {
    if (i < 1 || i > L - 3) {
#pragma unroll
        for (int q = 0; q < 9; ++q) O[q] = 0.f;
        return;
    }
    V3 p0 = load3(X, i - 1, L), p1 = load3(X, i, L), p2 = load3(X, i + 1, L);
    V3 u2 = gated_unit_bond(p0, p1), u1 = gated_unit_bond(p1, p2);
    V3 n2 = unit(cross(u2, u1));
    V3 o1 = unit(sub(u2, u1));
    V3 t = cross(o1, n2);
    O[0] = o1.x; O[1] = o1.y; O[2] = o1.z;
    O[3] = n2.x; O[4] = n2.y; O[5] = n2.z;
    O[6] = t.x;  O[7] = t.y;  O[8] = t.z;
}

__device__ __forceinline__ float pair_dist(V3 a, V3 b) {
    V3 d = sub(a, b);
    return sqrtf(dot(d, d) + 1e-6f);
}

template <bool OUT_F16>
__global__ void __launch_bounds__(256) edge_features_kernel(
    const float* __restrict__ X, const int* __restrict__ nbr_idx, const float* __restrict__ nbr_dist, int L, int K,
    const float* __restrict__ pos_table, const float* __restrict__ wedge_t, const float* __restrict__ ln_w,
    const float* __restrict__ ln_b, const float* __restrict__ we_t, const float* __restrict__ we_b,
    float* __restrict__ E_dbg, void* __restrict__ hE0, float* __restrict__ raw_out) {
    extern __shared__ __align__(16) float smem[];
    float* sRaw = smem;                       // [64][152]
    float* sE = sRaw + MAXK * RAW_LD;         // [64][128]
    int* sJ = reinterpret_cast<int*>(sE + MAXK * 128);   // [64]
    float* sOi = reinterpret_cast<float*>(sJ + MAXK);    // [9]

    const int f = blockIdx.y, i = blockIdx.x;
    const float* Xf = X + (size_t)f * L * 3;
    const size_t node = (size_t)f * L + i;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid < MAXK) sJ[tid] = tid < K ? nbr_idx[node * K + tid] : 0;
    if (tid == 0) residue_frame(Xf, i, L, sOi);
    for (int t = tid; t < MAXK * RAW_LD; t += 256) sRaw[t] = 0.f;   // rows >= K and the pad column stay zero
    __syncthreads();

    // ---- phase A: raw features, one warp per row ----
    const V3 ti[3] = {load3(Xf, i - 1, L), load3(Xf, i, L), load3(Xf, i + 1, L)};
    for (int r = warp; r < K; r += 8) {
        const int j = sJ[r];
        const V3 tj[3] = {load3(Xf, j - 1, L), load3(Xf, j, L), load3(Xf, j + 1, L)};
        for (int t = lane; t < 144; t += 32) {
            const int pr = t >> 4, m = t & 15;
            float D;
            switch (pr) {   // order of RBF_all, protein_mpnn_utils.py:493-505
                case 0: D = nbr_dist[node * K + r]; break;
                case 1: D = pair_dist(ti[0], tj[0]); break;
                case 2: D = pair_dist(ti[2], tj[2]); break;
                case 3: D = pair_dist(ti[0], tj[1]); break;
                case 4: D = pair_dist(ti[0], tj[2]); break;
                case 5: D = pair_dist(ti[1], tj[0]); break;
                case 6: D = pair_dist(ti[1], tj[2]); break;
                case 7: D = pair_dist(ti[2], tj[0]); break;
                default: D = pair_dist(ti[2], tj[1]); break;
            }
            const float z = (D - c_rbf_mu[m]) / 1.25f;
            sRaw[r * RAW_LD + t] = expf(-(z * z));
        }
        if (lane == 0) {
            float Oj[9];
            residue_frame(Xf, j, L, Oj);
            const V3 d = sub(tj[1], ti[1]);
            V3 dU = {dot({sOi[0], sOi[1], sOi[2]}, d), dot({sOi[3], sOi[4], sOi[5]}, d), dot({sOi[6], sOi[7], sOi[8]}, d)};
            dU = unit(dU);
            float R[3][3];   // R = O_i^T O_j
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) R[a][c] = (sOi[a] * Oj[c] + sOi[3 + a] * Oj[3 + c]) + sOi[6 + a] * Oj[6 + c];
            const float Rxx = R[0][0], Ryy = R[1][1], Rzz = R[2][2];
            auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
            float qx = sgn(R[2][1] - R[1][2]) * 0.5f * sqrtf(fabsf(1.f + (Rxx - Ryy - Rzz)));
            float qy = sgn(R[0][2] - R[2][0]) * 0.5f * sqrtf(fabsf(1.f + (-Rxx + Ryy - Rzz)));
            float qz = sgn(R[1][0] - R[0][1]) * 0.5f * sqrtf(fabsf(1.f + (-Rxx - Ryy + Rzz)));
            float qw = sqrtf(fmaxf(1.f + ((Rxx + Ryy) + Rzz), 0.f)) / 2.f;
            float qn = fmaxf(sqrtf((qx * qx + qy * qy) + (qz * qz + qw * qw)), 1e-12f);
            float* o = sRaw + r * RAW_LD + 144;
            o[0] = dU.x; o[1] = dU.y; o[2] = dU.z;
            o[3] = qx / qn; o[4] = qy / qn; o[5] = qz / qn; o[6] = qw / qn;
        }
    }
    __syncthreads();
    if (raw_out != nullptr) {
        // training step (train_ops.cu): the raw 144 RBF + 7 orientation features of every edge, [.., 152] (pad column zero); the
        // trainable projections are applied -- and differentiated -- as separate GEMMs
        for (int t = tid; t < K * RAW_LD; t += 256) raw_out[node * K * RAW_LD + t] = sRaw[t];
        if (hE0 == nullptr) return;
    }

    // ---- phase B: 151 -> 128 projection + positional table ----
    const int c = tid & 127, row0 = (tid >> 7) * 32;
    float acc[32];
    zero_acc(acc);
    tile_gemm<32, RAW_LD, RAW_LD>(sRaw + row0 * RAW_LD, wedge_t + c, 128, acc);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const int row = row0 + r;
        int d = i - sJ[row] + 32;
        d = d < 0 ? 0 : (d > 64 ? 64 : d);
        sE[row * 128 + c] = acc[r] + pos_table[d * 128 + c];
    }
    __syncthreads();

    // ---- phase C: LayerNorm (affine, eps 1e-5), then W_e ----
    for (int r = warp; r < MAXK; r += 8) {
        float4 v4 = *reinterpret_cast<float4*>(sE + r * 128 + lane * 4);
        float v[4] = {v4.x, v4.y, v4.z, v4.w};
        float mean, rstd;
        warp_ln_stats(v, 1e-5f, mean, rstd);
        const float4 w4 = *reinterpret_cast<const float4*>(ln_w + lane * 4);
        const float4 b4 = *reinterpret_cast<const float4*>(ln_b + lane * 4);
        float4 o;
        o.x = (v[0] - mean) * rstd * w4.x + b4.x;
        o.y = (v[1] - mean) * rstd * w4.y + b4.y;
        o.z = (v[2] - mean) * rstd * w4.z + b4.z;
        o.w = (v[3] - mean) * rstd * w4.w + b4.w;
        *reinterpret_cast<float4*>(sE + r * 128 + lane * 4) = o;
        if (E_dbg != nullptr && r < K) *reinterpret_cast<float4*>(E_dbg + (node * K + r) * 128 + lane * 4) = o;
    }
    __syncthreads();
    zero_acc(acc);
    tile_gemm<32, 128, 128>(sE + row0 * 128, we_t + c, 128, acc);
    const float bias = we_b[c];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const int row = row0 + r;
        if (row < K) {
            const float v = acc[r] + bias;
            if (OUT_F16) reinterpret_cast<__half*>(hE0)[(node * K + row) * 128 + c] = __float2half_rn(v);
            else reinterpret_cast<float*>(hE0)[(node * K + row) * 128 + c] = v;
        }
    }
}

}  // namespace

int launch_edge_features(const DenoiserModel& m, const float* X, const int* lengths, const int* idx, const float* D,
                         int F, int L, int K, float* E_dbg, void* hE0, int precision, cudaStream_t s) {
    (void)lengths;   // padding is expressed by zero coordinates beyond each frame's length, as in the reference
    if (K > MAXK) { set_error("edge_features: K=%d > %d", K, MAXK); return (int)cudaErrorInvalidValue; }
    static bool mu_ready = false;
    if (!mu_ready) {
        float mu[16];
        const float step = (22.0f - 2.0f) / 15.0f;     // torch.linspace(2, 22, 16), float arithmetic
        for (int q = 0; q < 16; ++q) mu[q] = q < 8 ? 2.0f + step * q : 22.0f - step * (15 - q);
        CB2_CUDA(cudaMemcpyToSymbol(c_rbf_mu, mu, sizeof(mu)));
        mu_ready = true;
    }
    const size_t smem = (size_t)(MAXK * RAW_LD + MAXK * 128) * 4 + MAXK * 4 + 16 * 4;
    dim3 grid(L, F);
    if (precision == PREC_F16) {
        CB2_CUDA(cudaFuncSetAttribute(edge_features_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        edge_features_kernel<true><<<grid, 256, smem, s>>>(X, idx, D, L, K, m.pos_table, m.wedge_t, m.ln_w, m.ln_b, m.we_t, m.we_b, E_dbg, hE0, nullptr);
    } else {
        CB2_CUDA(cudaFuncSetAttribute(edge_features_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        edge_features_kernel<false><<<grid, 256, smem, s>>>(X, idx, D, L, K, m.pos_table, m.wedge_t, m.ln_w, m.ln_b, m.we_t, m.we_b, E_dbg, hE0, nullptr);
    }
    CB2_LAUNCH_CHECK();
    return 0;
}

int launch_edge_raw_features(const float* X, const int* idx, const float* D, int F, int L, int K, float* raw_out, cudaStream_t s) {
    if (K > MAXK) { set_error("edge_raw_features: K=%d > %d", K, MAXK); return (int)cudaErrorInvalidValue; }
    static bool mu_ready = false;
    if (!mu_ready) {
        float mu[16];
        const float step = (22.0f - 2.0f) / 15.0f;
        for (int q = 0; q < 16; ++q) mu[q] = q < 8 ? 2.0f + step * q : 22.0f - step * (15 - q);
        CB2_CUDA(cudaMemcpyToSymbol(c_rbf_mu, mu, sizeof(mu)));
        mu_ready = true;
    }
    const size_t smem = (size_t)(MAXK * RAW_LD + MAXK * 128) * 4 + MAXK * 4 + 16 * 4;
    CB2_CUDA(cudaFuncSetAttribute(edge_features_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_features_kernel<false><<<dim3(L, F), 256, smem, s>>>(X, idx, D, L, K, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, raw_out);
    CB2_LAUNCH_CHECK();
    return 0;
}

}  // namespace cb2
