// fp32 (SIMT) tier of the per-edge message MLPs -- the parity tier: every product and sum is fp32,
// so results track the CPU oracle to ~1e-6 and validate the data flow that the tcgen05 tier
// (edge_tc.cu) shares.  One CTA per (member, residue) = its K <= 64 neighbour rows; 256 threads,
// thread = (output column, 32-row half) as in tile_gemm.cuh.
//
//   ENC_NODE  m = gelu(W2 gelu(W1b h_E + Pa[i] + Pc[j]) + b2);  S[i] = sum_k mask_ik m      (W3 is applied
//             per node in node.cu; reference protein_mpnn_utils.py:240-247)
//   ENC_EDGE  h_E' = gate3 * modulate(LN(h_E + W13 gelu(W12 gelu(W11b h_E + Pa'[i] + Pc'[j]) + b12) + b13))
//             (reference :261-270)
//   DEC       like ENC_NODE with the decoder's W1 blocks and NO neighbour mask (reference :300-307,
//             latent_model.py:258-262: h_ESV + h_EXV_encoder = [2 h_E | 2 h_S_j | h_V_j + h_Venc_j])
// Pa / Pc are the per-node halves of the first layer produced by node.cu; the 384/512-wide concat the
// reference materialises never exists.
#include "model.h"
#include "tile_gemm.cuh"

namespace cb2 {

namespace {

constexpr int MAXK = 64;

struct EdgeParams {
    int mode, L, K;
    const float* hE_in;      // rows of this node at hE_in + ((src_member*L + i)*K)*128 ; src_member = frame (layer 0) or b
    int in_is_frame;         // 1: index hE_in by frame_of[b]
    float* hE_out;           // ENC_EDGE
    const float* P;          // [N,256]
    const float *Wx_t, *W2_t, *b2, *W3_t, *b3;   // Wx = first-layer h_E block; W3 only for ENC_EDGE (W13)
    const float* mod;        // ENC_EDGE: this layer's adaLN block; row b at mod + b*mod_stride
    int mod_stride;
    const int *lengths, *frame_of, *nbr_idx;
    float* S;                // [N,128]
};

__global__ void __launch_bounds__(256) edge_f32_kernel(EdgeParams p) {
    extern __shared__ __align__(16) float smem[];
    float* sA = smem;                // [64][128]
    float* sB = sA + MAXK * 128;     // [64][128]
    float* sRed = sB + MAXK * 128;   // [128]
    int* sJ = reinterpret_cast<int*>(sRed + 128);   // [64]
    float* sMk = reinterpret_cast<float*>(sJ + MAXK);   // [64]

    const int tid = threadIdx.x, c = tid & 127, half = tid >> 7, row0 = half * 32;
    const int i = blockIdx.x, b = blockIdx.y;
    const int f = p.frame_of[b];
    const int len = p.lengths[f];
    const size_t n = (size_t)b * p.L + i;
    const int K = p.K;

    if (tid < MAXK) {
        int j = 0;
        float mk = 0.f;
        if (tid < K) {
            j = p.nbr_idx[((size_t)f * p.L + i) * K + tid];
            mk = (p.mode == EDGE_DEC) ? 1.f : ((i < len && j < len) ? 1.f : 0.f);
        }
        sJ[tid] = j; sMk[tid] = mk;
    }
    {
        const size_t src = ((size_t)(p.in_is_frame ? f : b) * p.L + i) * K * 128;
        const float4* g = reinterpret_cast<const float4*>(p.hE_in + src);
        for (int t = tid; t < MAXK * 32; t += 256) {
            const int r = t >> 5;
            reinterpret_cast<float4*>(sA)[t] = r < K ? g[t] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();

    float acc[32];
    // ---- layer 1: W?b h_E + per-node halves ----
    zero_acc(acc);
    tile_gemm<32, 128, 128>(sA + row0 * 128, p.Wx_t + c, 128, acc);
    {
        const float pa = p.P[n * 256 + c];
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int row = row0 + r;
            const float pc = p.P[((size_t)b * p.L + sJ[row]) * 256 + 128 + c];
            sB[row * 128 + c] = gelu_erf((acc[r] + pa) + pc);
        }
    }
    __syncthreads();
    // ---- layer 2 ----
    zero_acc(acc);
    tile_gemm<32, 128, 128>(sB + row0 * 128, p.W2_t + c, 128, acc);
    const float b2 = p.b2[c];
    if (p.mode != EDGE_ENC_EDGE) {
        float part = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int row = row0 + r;
            if (row < K) part += sMk[row] * gelu_erf(acc[r] + b2);
        }
        if (half == 1) sRed[c] = part;
        __syncthreads();
        if (half == 0) p.S[n * 128 + c] = part + sRed[c];
        return;
    }
    __syncthreads();                 // all reads of sB (layer-2 input) are done
#pragma unroll
    for (int r = 0; r < 32; ++r) sB[(row0 + r) * 128 + c] = gelu_erf(acc[r] + b2);
    __syncthreads();
    // ---- layer 3 + residual + LayerNorm + adaLN (edge stream) ----
    zero_acc(acc);
    tile_gemm<32, 128, 128>(sB + row0 * 128, p.W3_t + c, 128, acc);
    const float b3 = p.b3[c];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const int row = row0 + r;
        sB[row * 128 + c] = sA[row * 128 + c] + (acc[r] + b3);
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const float* m = p.mod + (size_t)b * p.mod_stride;      // [.. shift3 @768 | scale3 @896 | gate3 @1024]
    const float4 sh = *reinterpret_cast<const float4*>(m + 768 + lane * 4);
    const float4 sc = *reinterpret_cast<const float4*>(m + 896 + lane * 4);
    const float4 gt = *reinterpret_cast<const float4*>(m + 1024 + lane * 4);
    float* out = p.hE_out + n * K * 128;
    for (int r = warp; r < K; r += 8) {
        float4 v4 = *reinterpret_cast<float4*>(sB + r * 128 + lane * 4);
        float v[4] = {v4.x, v4.y, v4.z, v4.w};
        float mean, rstd;
        warp_ln_stats(v, 1e-6f, mean, rstd);
        float4 o;
        o.x = gt.x * ((v[0] - mean) * rstd * (1.0f + sc.x) + sh.x);
        o.y = gt.y * ((v[1] - mean) * rstd * (1.0f + sc.y) + sh.y);
        o.z = gt.z * ((v[2] - mean) * rstd * (1.0f + sc.z) + sh.z);
        o.w = gt.w * ((v[3] - mean) * rstd * (1.0f + sc.w) + sh.w);
        *reinterpret_cast<float4*>(out + r * 128 + lane * 4) = o;
    }
}

}  // namespace

int launch_edge_f32(Plan& p, int mode, int layer, const float* mod_base, int mod_stride_b, cudaStream_t s) {
    const DenoiserModel& m = *p.model;
    EdgeParams ep{};
    ep.mode = mode; ep.L = p.L; ep.K = p.K;
    ep.lengths = p.lengths; ep.frame_of = p.frame_of; ep.nbr_idx = p.nbr_idx; ep.S = p.S;
    ep.mod_stride = mod_stride_b;
    const bool first = (layer == 0 && mode != EDGE_DEC);
    ep.hE_in = reinterpret_cast<const float*>(first ? p.hE0 : p.hE);
    ep.in_is_frame = first ? 1 : 0;
    if (mode == EDGE_ENC_NODE) {
        const EncLayerW& e = m.enc[layer];
        ep.P = plan_P(p, 0); ep.Wx_t = e.W1b_t; ep.W2_t = e.W2_t; ep.b2 = e.b2;
    } else if (mode == EDGE_ENC_EDGE) {
        const EncLayerW& e = m.enc[layer];
        ep.P = plan_P(p, 1); ep.Wx_t = e.W11b_t; ep.W2_t = e.W12_t; ep.b2 = e.b12; ep.W3_t = e.W13_t; ep.b3 = e.b13;
        ep.hE_out = reinterpret_cast<float*>(p.hE);
        ep.mod = mod_base + CB2_MOD_ENC_OFF(layer);
    } else {
        const DecLayerW& d = m.dec[layer];
        ep.P = plan_P(p, 0); ep.Wx_t = d.W1b2_t; ep.W2_t = d.W2_t; ep.b2 = d.b2;
    }
    const size_t smem = (size_t)(2 * MAXK * 128 + 128) * 4 + MAXK * 8;
    static bool attr = false;
    if (!attr) {
        CB2_CUDA(cudaFuncSetAttribute(edge_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    edge_f32_kernel<<<dim3(p.L, p.NB), 256, smem, s>>>(ep);
    CB2_LAUNCH_CHECK();
    p.launches++;
    return 0;
}

}  // namespace cb2
