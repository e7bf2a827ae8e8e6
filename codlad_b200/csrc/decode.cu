// Decode side of the sampling path: de-normalise -> VQ codebook lookup -> map_out -> IC decoder ->
// internal coordinates -> Cartesian.
//
//  * vq_lookup      get_norm_feature(norm_in=False) (reference utils/dataset_module.py:230-256) fused with
//                   the eval-mode nearest-code search of vector_quantize_pytorch==1.21.7 (third party;
//                   in-repo equivalent utils/vq_module.py:61-68) and map_out (models/vae_model.py:761-762).
//                   Bit-exact against oracle/restate.py::vq_nearest: expanded |x|^2+|e|^2-2x.e with the
//                   oracle's association order, every op individually rounded (no FMA), clamp, sqrt,
//                   first index on ties.
//  * ic_decoder_*   IC_Decoder / IC_Decoder_angle.forward (models/vae_model.py:375-412,467-503) with
//                   InvariantMessage / DistanceEmbed (models/gcn_nn.py:222-381).  The radius graph is a
//                   per-frame CSR shared by all ensemble members; scatter_add becomes a deterministic
//                   segmented sum (no atomics); the distance filter w = Linear(sincRBF15(d)) * cosine
//                   envelope depends only on the frame, so it is tabulated once per frame per block.
//  * ic_to_xyz      utils/utils_ic.py:197-268: 13 sequential atom placements per residue held in
//                   registers, one thread per (member, residue); the reference's final
//                   `reshape[:, atom_idx][:, permute]` compaction is a precomputed slot -> atom map.
#include "decode.h"

namespace cb2 {

namespace {

// ------------------------------------------------------------------ VQ
__global__ void codebook_norms_kernel(const float* __restrict__ cb, int M, float* __restrict__ e2) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float a = cb[m * 3], b = cb[m * 3 + 1], c = cb[m * 3 + 2];
    e2[m] = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
}

constexpr int VQ_WARPS = 8;         // warps per CTA
constexpr int VQ_PER_WARP = 4;      // residues per warp (the codebook is staged once per CTA for 32 residues)

// One warp per residue: lanes stride the codes (each lane keeps the first minimum of its own increasing subsequence), then a
// lexicographic (distance, index) minimum across the lanes -- the index of the first minimum, exactly what a sequential scan
// (and the reference's argmax over -dist) returns.  The distance arithmetic per code is unchanged.
__global__ void __launch_bounds__(VQ_WARPS * 32) vq_lookup_kernel(const float* __restrict__ x, int N, int L, const int* __restrict__ lengths,
                                                                  const int* __restrict__ frame_of, const float* __restrict__ mean,
                                                                  const float* __restrict__ stdv, int denorm, const float* __restrict__ cb,
                                                                  const float* __restrict__ e2, int M, const float* __restrict__ mapw_t,
                                                                  const float* __restrict__ mapb, int* __restrict__ idx_out,
                                                                  float* __restrict__ zq_out, float* __restrict__ S40) {
    extern __shared__ __align__(16) float smem[];
    float* sCb = smem;            // [M*3]
    float* sE2 = smem + M * 3;    // [M]
    for (int t = threadIdx.x; t < M * 3; t += blockDim.x) sCb[t] = cb[t];
    for (int t = threadIdx.x; t < M; t += blockDim.x) sE2[t] = e2[t];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = 0; k < VQ_PER_WARP; ++k) {
        const int n = (blockIdx.x * VQ_WARPS + warp) * VQ_PER_WARP + k;
        if (n >= N) return;
        float v0 = x[n * 3], v1 = x[n * 3 + 1], v2 = x[n * 3 + 2];
        if (denorm) {   // feature * std + mean, two roundings like the reference's torch ops
            v0 = __fadd_rn(__fmul_rn(v0, stdv[0]), mean[0]);
            v1 = __fadd_rn(__fmul_rn(v1, stdv[1]), mean[1]);
            v2 = __fadd_rn(__fmul_rn(v2, stdv[2]), mean[2]);
        }
        const int b = n / L, i = n - b * L;
        const bool valid = i < lengths[frame_of[b]];
        int best = -1;
        float q0 = v0, q1 = v1, q2 = v2;
        if (valid) {            // (warp-uniform: every lane works on the same residue)
            const float x2 = __fadd_rn(__fadd_rn(__fmul_rn(v0, v0), __fmul_rn(v1, v1)), __fmul_rn(v2, v2));
            float bestd = 0.f;
            for (int m = lane; m < M; m += 32) {
                const float dotp = __fadd_rn(__fadd_rn(__fmul_rn(v0, sCb[m * 3]), __fmul_rn(v1, sCb[m * 3 + 1])), __fmul_rn(v2, sCb[m * 3 + 2]));
                const float s = __fadd_rn(__fadd_rn(x2, sE2[m]), __fmul_rn(dotp, -2.0f));
                const float d = __fsqrt_rn(fmaxf(s, 0.0f));
                if (best < 0 || d < bestd) { best = m; bestd = d; }
            }
            for (int off = 16; off >= 1; off >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bestd, off);
                const int ob = __shfl_xor_sync(0xffffffffu, best, off);
                if (ob >= 0 && (best < 0 || od < bestd || (od == bestd && ob < best))) { best = ob; bestd = od; }
            }
            q0 = sCb[best * 3]; q1 = sCb[best * 3 + 1]; q2 = sCb[best * 3 + 2];
        }
        if (lane == 0) {
            idx_out[n] = best;
            zq_out[n * 3] = q0; zq_out[n * 3 + 1] = q1; zq_out[n * 3 + 2] = q2;
        }
        if (S40 != nullptr) {   // map_out: Linear(3 -> 36); columns 36..39 (residue embedding) are filled by the decoder
            for (int o = lane; o < 36; o += 32)
                S40[(size_t)n * 40 + o] = fmaf(q2, mapw_t[2 * 36 + o], fmaf(q1, mapw_t[36 + o], fmaf(q0, mapw_t[o], mapb[o])));
        }
    }
}

// ------------------------------------------------------------------ IC decoder
__device__ __forceinline__ float swishf(float x) { return x / (1.0f + expf(-x)); }   // x * sigmoid(x)

// Distance filters of the 4 message blocks for every directed frame edge: w[blk][e][40].
__global__ void __launch_bounds__(128) ic_edge_filter_kernel(const float* __restrict__ X, int L, const int* __restrict__ row_ptr,
                                                             const int* __restrict__ col, int n_rows, int E,
                                                             const float* __restrict__ Wd_t /* [4][15][40] */,
                                                             const float* __restrict__ bd /* [4][40] */, float cutoff,
                                                             float* __restrict__ w /* [4][E][40] */) {
    // one CTA per source row (frame node); its warps stride the row's edges.  Per edge the 15 sinc basis functions and the cosine
    // envelope are evaluated once (lane q < 15: basis q, lane 15: envelope) and broadcast; the lanes then stride the 4 x 40 outputs.
    const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const int f = row / L;
    const float* Xf = X + (size_t)f * L * 3;
    const int i = row - f * L;
    const float xi = Xf[i * 3], yi = Xf[i * 3 + 1], zi = Xf[i * 3 + 2];
    const float kPi = 3.14159265358979323846f;
    for (int e = row_ptr[row] + warp; e < row_ptr[row + 1]; e += blockDim.x >> 5) {
        const int j = col[e];
        const float dx = Xf[j * 3] - xi, dy = Xf[j * 3 + 1] - yi, dz = Xf[j * 3 + 2] - zi;
        const float dist = sqrtf(((dx * dx + 1e-8f) + (dy * dy + 1e-8f)) + (dz * dz + 1e-8f));   // preprocess_r, gcn_nn.py:66-70
        float mine = 0.f;
        if (lane < 15) {
            const float coef = (float)(lane + 1) * kPi / cutoff;
            mine = dist >= cutoff ? 0.f : (dist == 0.f ? coef : sinf(coef * dist) / dist);
        } else if (lane == 15) {
            mine = dist >= cutoff ? 0.f : 0.5f * (cosf(kPi * dist / cutoff) + 1.0f);
        }
        float rbf[15];
#pragma unroll
        for (int q = 0; q < 15; ++q) rbf[q] = __shfl_sync(0xffffffffu, mine, q);
        const float env = __shfl_sync(0xffffffffu, mine, 15);
        for (int t = lane; t < 160; t += 32) {
            const int blk = t / 40, c = t - blk * 40;
            float a = bd[blk * 40 + c];
#pragma unroll
            for (int q = 0; q < 15; ++q) a = fmaf(rbf[q], Wd_t[(blk * 15 + q) * 40 + c], a);
            w[((size_t)blk * E + e) * 40 + c] = a * env;
        }
    }
}

// out[c] = b[c] + sum_k in[k] W_t[k][c]  for c < n_out; `in` is a shared-memory vector.
__device__ __forceinline__ float matvec(const float* in, const float* __restrict__ W_t, const float* __restrict__ b, int n_in,
                                        int n_out, int c) {
    float a = b[c];
    for (int k = 0; k < n_in; ++k) a = fmaf(in[k], W_t[k * n_out + c], a);
    return a;
}

constexpr int ICT = 64;   // threads per node (>= 50 features)

// S[:, 36:40] = res_embed(cg_z)  (vae_model.py:477 / :385)
__global__ void ic_embed_kernel(float* __restrict__ S40, const int* __restrict__ cg_z, const int* __restrict__ frame_of, int L, int N,
                                const float* __restrict__ res_embed) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * 4) return;
    const int n = t >> 2, q = t & 3;
    const int b = n / L, i = n - b * L;
    S40[(size_t)n * 40 + 36 + q] = res_embed[cg_z[(size_t)frame_of[b] * L + i] * 4 + q];
}

// phi = Dense(swish(Dense(S)))   (InvariantMessage.inv_dense, gcn_nn.py:351-359)
__global__ void __launch_bounds__(256) ic_phi_kernel(const float* __restrict__ S40, int N, const float* __restrict__ W0_t,
                                                     const float* __restrict__ b0, const float* __restrict__ W1_t,
                                                     const float* __restrict__ b1, float* __restrict__ phi) {
    __shared__ float sIn[4][40], sMid[4][40];
    const int g = threadIdx.x / ICT, c = threadIdx.x % ICT;
    const int n = blockIdx.x * 4 + g;
    if (n < N && c < 40) sIn[g][c] = S40[(size_t)n * 40 + c];
    __syncthreads();
    if (n < N && c < 40) sMid[g][c] = swishf(matvec(sIn[g], W0_t, b0, 40, 40, c));
    __syncthreads();
    if (n < N && c < 40) phi[(size_t)n * 40 + c] = matvec(sMid[g], W1_t, b1, 40, 40, c);
}

// v_i = sum_j phi_j * w_ij ;  S += Linear(swish(Linear(swish(v))))   (vae_model.py:480-490)
__global__ void __launch_bounds__(256) ic_message_kernel(float* __restrict__ S40, const float* __restrict__ phi, int N, int L,
                                                         const int* __restrict__ frame_of, const int* __restrict__ lengths,
                                                         const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                         const float* __restrict__ w /* [E][40] of this block */,
                                                         const float* __restrict__ W1_t, const float* __restrict__ b1,
                                                         const float* __restrict__ W3_t, const float* __restrict__ b3) {
    __shared__ float sV[4][40], sMid[4][40];
    const int g = threadIdx.x / ICT, c = threadIdx.x % ICT;
    const int n = blockIdx.x * 4 + g;
    const bool on = n < N && c < 40;
    if (on) {
        const int b = n / L, i = n - b * L, f = frame_of[b];
        float v = 0.f;
        if (i < lengths[f]) {
            const int r = f * L + i;
            for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e)
                v = fmaf(phi[((size_t)b * L + col[e]) * 40 + c], w[(size_t)e * 40 + c], v);
        }
        sV[g][c] = swishf(v);
    }
    __syncthreads();
    if (on) sMid[g][c] = swishf(matvec(sV[g], W1_t, b1, 40, 40, c));
    __syncthreads();
    if (on) S40[(size_t)n * 40 + c] += matvec(sMid[g], W3_t, b3, 40, 40, c);
}

struct HeadW {
    const float *bb_dist, *sc_dist;                   // [25][3], [25][10]
    const float *ba1_t, *ba1_b, *ba3_t, *ba3_b;       // backbone_angle 40->3->3
    const float *bt1_t, *bt1_b, *bt3_t, *bt3_b;       // backbone_torsion 43->3->3
    const float *sa_embed;                            // N6: [25][10]
    const float *sa1_t, *sa1_b, *sa3_t, *sa3_b;       // angle variant 40->10->10
    const float *tb1_t[4], *tb1_b[4], *tb3_t[4], *tb3_b[4];   // sidechain_torsion_blocks T->T->T
    const float *ft1_t, *ft1_b, *ft3_t, *ft3_b;       // final_torsion T->10->10
    int angle_variant;
};

// All heads of the decoder for one node -> ic_recon[n][13][3] = (bond, angle, torsion)  (vae_model.py:491-503 / :398-412)
__global__ void __launch_bounds__(256) ic_heads_kernel(const float* __restrict__ S40, int N, int L, const int* __restrict__ frame_of,
                                                       const int* __restrict__ cg_z, HeadW hw, float* __restrict__ ic) {
    __shared__ float sS[4][56], sT[4][56], sU[4][56], sSmall[4][16];
    const int g = threadIdx.x / ICT, c = threadIdx.x % ICT;
    const int n = blockIdx.x * 4 + g;
    const bool live = n < N;
    const int T = hw.angle_variant ? 50 : 40;
    int z = 0;
    if (live) {
        const int b = n / L, i = n - b * L;
        z = cg_z[(size_t)frame_of[b] * L + i];
        if (c < 40) sS[g][c] = S40[(size_t)n * 40 + c];
    }
    __syncthreads();
    float* out = ic + (size_t)n * 39;
    // backbone angle: Sequential(swish, Linear(40,3), swish, Linear(3,3))
    if (live && c < 40) sT[g][c] = swishf(sS[g][c]);
    __syncthreads();
    if (live && c < 3) sSmall[g][c] = swishf(matvec(sT[g], hw.ba1_t, hw.ba1_b, 40, 3, c));
    __syncthreads();
    if (live && c < 3) sSmall[g][4 + c] = matvec(sSmall[g], hw.ba3_t, hw.ba3_b, 3, 3, c);     // bb_angle
    __syncthreads();
    // backbone torsion on cat([S, bb_angle]) (43)
    if (live && c < 3) sT[g][40 + c] = swishf(sSmall[g][4 + c]);
    __syncthreads();
    if (live && c < 3) sSmall[g][8 + c] = swishf(matvec(sT[g], hw.bt1_t, hw.bt1_b, 43, 3, c));
    __syncthreads();
    if (live && c < 3) {
        const float tors = matvec(sSmall[g] + 8, hw.bt3_t, hw.bt3_b, 3, 3, c);
        out[c * 3 + 0] = hw.bb_dist[z * 3 + c];
        out[c * 3 + 1] = sSmall[g][4 + c];
        out[c * 3 + 2] = tors;
    }
    __syncthreads();
    // side-chain angle
    if (hw.angle_variant) {
        if (live && c < 40) sT[g][c] = swishf(sS[g][c]);
        __syncthreads();
        if (live && c < 10) sU[g][c] = swishf(matvec(sT[g], hw.sa1_t, hw.sa1_b, 40, 10, c));
        __syncthreads();
        if (live && c < 10) sS[g][40 + c] = matvec(sU[g], hw.sa3_t, hw.sa3_b, 10, 10, c);      // sc_S = cat([S, sc_angle])
        __syncthreads();
    }
    if (live && c < 10) {
        out[(3 + c) * 3 + 0] = hw.sc_dist[z * 10 + c];
        out[(3 + c) * 3 + 1] = hw.angle_variant ? sS[g][40 + c] : hw.sa_embed[z * 10 + c];
    }
    // residual torsion blocks
    for (int blk = 0; blk < 4; ++blk) {
        if (live && c < T) sT[g][c] = swishf(sS[g][c]);
        __syncthreads();
        if (live && c < T) sU[g][c] = swishf(matvec(sT[g], hw.tb1_t[blk], hw.tb1_b[blk], T, T, c));
        __syncthreads();
        if (live && c < T) sS[g][c] += matvec(sU[g], hw.tb3_t[blk], hw.tb3_b[blk], T, T, c);
        __syncthreads();
    }
    if (live && c < T) sT[g][c] = swishf(sS[g][c]);
    __syncthreads();
    if (live && c < 10) sU[g][c] = swishf(matvec(sT[g], hw.ft1_t, hw.ft1_b, T, 10, c));
    __syncthreads();
    if (live && c < 10) out[(3 + c) * 3 + 2] = matvec(sU[g], hw.ft3_t, hw.ft3_b, 10, 10, c);
}

// ------------------------------------------------------------------ internal coordinates -> Cartesian
struct P3 { float x, y, z; };
__device__ __forceinline__ P3 psub(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float pdot(P3 a, P3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ P3 pcross(P3 a, P3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

__device__ __forceinline__ P3 rodrigues(P3 axis, float angle, P3 v) {   // utils_ic.py:197-210
    const float nrm = sqrtf(pdot(axis, axis));
    axis = {axis.x / nrm, axis.y / nrm, axis.z / nrm};
    const float a = cosf(angle / 2.0f), sn = sinf(angle / 2.0f);
    const float b = -axis.x * sn, c = -axis.y * sn, d = -axis.z * sn;
    const float r00 = a * a + b * b - c * c - d * d, r01 = 2.f * (b * c - a * d), r02 = 2.f * (b * d + a * c);
    const float r10 = 2.f * (b * c + a * d), r11 = a * a + c * c - b * b - d * d, r12 = 2.f * (c * d - a * b);
    const float r20 = 2.f * (b * d - a * c), r21 = 2.f * (c * d + a * b), r22 = a * a + d * d - b * b - c * c;
    return {(r00 * v.x + r01 * v.y) + r02 * v.z, (r10 * v.x + r11 * v.y) + r12 * v.z, (r20 * v.x + r21 * v.y) + r22 * v.z};
}

__device__ __forceinline__ P3 place_atom(const float* ic3, P3 p1, P3 p2, P3 p3) {   // utils_ic.py:213-239
    P3 a = psub(p2, p1), b = psub(p2, p3);
    if (a.x == 0.f) a.x = 1e-8f; if (a.y == 0.f) a.y = 1e-8f; if (a.z == 0.f) a.z = 1e-8f;
    if (b.x == 0.f) b.x = 1e-8f; if (b.y == 0.f) b.y = 1e-8f; if (b.z == 0.f) b.z = 1e-8f;
    const float bond = fabsf(ic3[0]), an = sqrtf(pdot(a, a));
    P3 d = {bond * a.x / an, bond * a.y / an, bond * a.z / an};
    d = rodrigues(pcross(a, b), ic3[1], d);
    d = rodrigues(a, ic3[2], d);
    return {p1.x + d.x, p1.y + d.y, p1.z + d.z};
}

__global__ void __launch_bounds__(128) ic_to_xyz_kernel(const float* __restrict__ ca_full /* [F][L+2][3] */,
                                                        const float* __restrict__ ic /* [N][13][3] */, int N, int L,
                                                        const int* __restrict__ frame_of, const int* __restrict__ lengths,
                                                        const signed char* __restrict__ orders /* [F][L][10][3] */,
                                                        const int* __restrict__ slot_atom /* [F][L*14] */,
                                                        const long long* __restrict__ out_off /* [NB] atom offset */,
                                                        float* __restrict__ xyz /* [sum Na][3] */, float* __restrict__ slots_dbg) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = n / L, i = n - b * L, f = frame_of[b];
    if (i >= lengths[f]) return;
    const float* ca = ca_full + ((size_t)f * (L + 2) + i) * 3;     // ca[0..2]=prev, [3..5]=this, [6..8]=next
    const P3 prv = {ca[0], ca[1], ca[2]}, cur = {ca[3], ca[4], ca[5]}, nxt = {ca[6], ca[7], ca[8]};
    const float* q = ic + (size_t)n * 39;
    P3 s[14];
#pragma unroll
    for (int t = 4; t < 14; ++t) s[t] = cur;
    s[1] = place_atom(q + 0, cur, prv, nxt);        // N
    s[2] = place_atom(q + 3, cur, nxt, prv);        // C
    s[0] = place_atom(q + 6, s[2], cur, s[1]);      // O
    s[3] = cur;                                     // CA
    const signed char* ord = orders + ((size_t)f * L + i) * 30;
    for (int a = 0; a < 10; ++a) {
        const int o3 = ord[a * 3 + 0], o2 = ord[a * 3 + 1], o1 = ord[a * 3 + 2];
        P3 p1 = s[0], p2 = s[0], p3 = s[0];
#pragma unroll
        for (int t = 0; t < 13; ++t) {       // register-resident select (no local-memory indexing)
            if (t == o1) p1 = s[t];
            if (t == o2) p2 = s[t];
            if (t == o3) p3 = s[t];
        }
        const P3 v = place_atom(q + 9 + a * 3, p1, p2, p3);
#pragma unroll
        for (int t = 4; t < 14; ++t) if (t == 4 + a) s[t] = v;
    }
    const int* sa = slot_atom + ((size_t)f * L + i) * 14;
    float* o = xyz + (size_t)out_off[b] * 3;
#pragma unroll
    for (int t = 0; t < 14; ++t) {
        const int at = sa[t];
        if (at >= 0) { o[(size_t)at * 3] = s[t].x; o[(size_t)at * 3 + 1] = s[t].y; o[(size_t)at * 3 + 2] = s[t].z; }
        if (slots_dbg != nullptr) {
            float* sd = slots_dbg + ((size_t)n * 14 + t) * 3;
            sd[0] = s[t].x; sd[1] = s[t].y; sd[2] = s[t].z;
        }
    }
}

}  // namespace

int launch_codebook_norms(const float* cb, int M, float* e2, cudaStream_t s) {
    codebook_norms_kernel<<<(M + 255) / 256, 256, 0, s>>>(cb, M, e2);
    CB2_LAUNCH_CHECK();
    return 0;
}

int launch_vq_lookup(const VaeModel& v, const float* x, int N, int L, const int* lengths, const int* frame_of, int denorm,
                     int* idx_out, float* zq_out, float* S40, cudaStream_t s) {
    const size_t smem = (size_t)v.M * 16;
    if (smem > 200 * 1024) { set_error("vq: codebook of %d entries does not fit shared memory", v.M); return (int)cudaErrorInvalidValue; }
    CB2_CUDA(cudaFuncSetAttribute(vq_lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_cta = VQ_WARPS * VQ_PER_WARP;
    vq_lookup_kernel<<<(N + per_cta - 1) / per_cta, VQ_WARPS * 32, smem, s>>>(x, N, L, lengths, frame_of, v.mean, v.stdv, denorm, v.codebook, v.e2, v.M,
                                                        v.mapw_t, v.mapb, idx_out, zq_out, S40);
    CB2_LAUNCH_CHECK();
    return 0;
}

int launch_ic_edge_filters(const VaeModel& v, const float* X, int F, int L, const int* row_ptr, const int* col, int E, float* w,
                           cudaStream_t s) {
    const int rows = F * L;
    if (E == 0) return 0;
    ic_edge_filter_kernel<<<rows, 128, 0, s>>>(X, L, row_ptr, col, rows, E, v.Wd_t, v.bd, v.cutoff, w);
    CB2_LAUNCH_CHECK();
    return 0;
}

int launch_ic_decoder(const VaeModel& v, float* S40, float* phi, int N, int L, const int* frame_of, const int* lengths,
                      const int* cg_z, const int* row_ptr, const int* col, int E, const float* w, float* ic, cudaStream_t s,
                      long long* launches) {
    const int nb = (N + 3) / 4;
    ic_embed_kernel<<<(N * 4 + 255) / 256, 256, 0, s>>>(S40, cg_z, frame_of, L, N, v.res_embed);
    CB2_LAUNCH_CHECK();
    for (int blk = 0; blk < 4; ++blk) {
        ic_phi_kernel<<<nb, 256, 0, s>>>(S40, N, v.inv0_t[blk], v.inv0_b[blk], v.inv1_t[blk], v.inv1_b[blk], phi);
        CB2_LAUNCH_CHECK();
        ic_message_kernel<<<nb, 256, 0, s>>>(S40, phi, N, L, frame_of, lengths, row_ptr, col, w + (size_t)blk * E * 40,
                                             v.db1_t[blk], v.db1_b[blk], v.db3_t[blk], v.db3_b[blk]);
        CB2_LAUNCH_CHECK();
    }
    HeadW hw{};
    hw.bb_dist = v.bb_dist; hw.sc_dist = v.sc_dist;
    hw.ba1_t = v.ba1_t; hw.ba1_b = v.ba1_b; hw.ba3_t = v.ba3_t; hw.ba3_b = v.ba3_b;
    hw.bt1_t = v.bt1_t; hw.bt1_b = v.bt1_b; hw.bt3_t = v.bt3_t; hw.bt3_b = v.bt3_b;
    hw.sa_embed = v.sa_embed; hw.sa1_t = v.sa1_t; hw.sa1_b = v.sa1_b; hw.sa3_t = v.sa3_t; hw.sa3_b = v.sa3_b;
    for (int q = 0; q < 4; ++q) { hw.tb1_t[q] = v.tb1_t[q]; hw.tb1_b[q] = v.tb1_b[q]; hw.tb3_t[q] = v.tb3_t[q]; hw.tb3_b[q] = v.tb3_b[q]; }
    hw.ft1_t = v.ft1_t; hw.ft1_b = v.ft1_b; hw.ft3_t = v.ft3_t; hw.ft3_b = v.ft3_b;
    hw.angle_variant = v.angle_variant;
    ic_heads_kernel<<<nb, 256, 0, s>>>(S40, N, L, frame_of, cg_z, hw, ic);
    CB2_LAUNCH_CHECK();
    if (launches) *launches += 10;
    return 0;
}

int launch_ic_to_xyz(const float* ca_full, const float* ic, int N, int L, const int* frame_of, const int* lengths,
                     const signed char* orders, const int* slot_atom, const long long* out_off, float* xyz, float* slots_dbg,
                     cudaStream_t s) {
    ic_to_xyz_kernel<<<(N + 127) / 128, 128, 0, s>>>(ca_full, ic, N, L, frame_of, lengths, orders, slot_atom, out_off, xyz, slots_dbg);
    CB2_LAUNCH_CHECK();
    return 0;
}

}  // namespace cb2
