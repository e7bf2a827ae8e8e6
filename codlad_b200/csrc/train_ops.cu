// Kernels of the train_latent step: everything of the denoiser's training forward / backward that is
// not a plain GEMM (train_gemm.cu), plus the optimiser.  fp32 throughout; every reduction runs in a fixed order (no atomics), so a
// step is bit-reproducible.  Hidden width H = 128 is a compile-time constant (one warp = one row, 4 columns per lane).
//
//   edge_combine_gelu   z = Z + Pa[i] + Pc[j] (+ b), y = GELU(z)      first layer of the message MLPs with W1 factored into its
//                                                                      h_V_i / h_E / h_V_j blocks (protein_mpnn_utils.py:240-243,
//                                                                      261-264, 300-303: cat_neighbors_nodes + W1 / W11)
//   edge_gather_bwd     dPa[i] = sum_k dz[i,k];  dPc[j] = sum over the edges that point at j (reverse CSR)
//   masked_sum          dh[i] = sum_k mask[i,k] m[i,k] / 30                                                   (:244-246, :304-306)
//   ln_mod              y = rowmask * gate * (LN(a + drop * b) * (1 + scale) + shift)     (:247-259, :266-270; latent_model.py:18-35;
//                       with gate = null, scale = w - 1, shift = b it is the affine LayerNorm of the featuriser, :521)
//   bias_gelu / gelu_bwd / silu, index_sum (embedding gradients), colsum (bias gradients), sumsq (gradient norm),
//   adamw_ema           torch.optim.AdamW step + update_ema (train_latent.py:252-261, utils/train_module.py:101-111) in one pass
#include "../../include/codlad_b200_train.h"
#include "model.h"
#include "train_ops.h"

namespace cb2 {
namespace train {

namespace {

constexpr int H = 128;

__device__ __forceinline__ float gelu_grad(float x) {       // d/dx [0.5 x (1 + erf(x / sqrt 2))]
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

__global__ void bias_gelu_fwd_kernel(float* __restrict__ Z, const float* __restrict__ bias, long long n, int cols, float* __restrict__ Y) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    float4 z = *reinterpret_cast<float4*>(Z + i);
    if (bias != nullptr) {
        const float4 b = *reinterpret_cast<const float4*>(bias + (i % cols));
        z.x += b.x; z.y += b.y; z.z += b.z; z.w += b.w;
        *reinterpret_cast<float4*>(Z + i) = z;
    }
    if (Y != nullptr) *reinterpret_cast<float4*>(Y + i) = make_float4(gelu_erf(z.x), gelu_erf(z.y), gelu_erf(z.z), gelu_erf(z.w));
}

__global__ void gelu_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ dY, long long n, float* __restrict__ dpre) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const float4 z = *reinterpret_cast<const float4*>(pre + i), g = *reinterpret_cast<const float4*>(dY + i);
    *reinterpret_cast<float4*>(dpre + i) = make_float4(g.x * gelu_grad(z.x), g.y * gelu_grad(z.y), g.z * gelu_grad(z.z), g.w * gelu_grad(z.w));
}

__global__ void bias_add_scalar_kernel(float* __restrict__ Z, const float* __restrict__ bias, long long n, int cols) {      // cols % 4 != 0 (the 6-wide output layer)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) Z[i] += bias[i % cols];
}

// mode 0: y = silu(x); mode 1: dx = dy * silu'(x); mode 2: out = a + b; mode 3: out = a * s (s scalar in `scale`); mode 4: out = a * b
__global__ void elementwise_kernel(int mode, const float* __restrict__ a, const float* __restrict__ b, float scale, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = a[i];
    if (mode == 0) out[i] = silu(x);
    else if (mode == 1) { const float sg = 1.0f / (1.0f + expf(-x)); out[i] = b[i] * (sg * (1.0f + x * (1.0f - sg))); }
    else if (mode == 2) out[i] = x + b[i];
    else if (mode == 3) out[i] = x * scale;
    else out[i] = x * b[i];
}

__device__ __forceinline__ float ew_op(int mode, float x, float y, float scale) {
    if (mode == 0) return silu(x);
    if (mode == 1) { const float sg = 1.0f / (1.0f + expf(-x)); return y * (sg * (1.0f + x * (1.0f - sg))); }
    if (mode == 2) return x + y;
    if (mode == 3) return x * scale;
    return x * y;
}
__global__ void elementwise_vec_kernel(int mode, const float* __restrict__ a, const float* __restrict__ b, float scale, long long n4, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 x = reinterpret_cast<const float4*>(a)[i];
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mode == 1 || mode == 2 || mode == 4) y = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(out)[i] = make_float4(ew_op(mode, x.x, y.x, scale), ew_op(mode, x.y, y.y, scale), ew_op(mode, x.z, y.z, scale), ew_op(mode, x.w, y.w, scale));
}

// one warp per edge row
__global__ void __launch_bounds__(256) edge_combine_gelu_kernel(float* __restrict__ Z, const float* __restrict__ Pa, const float* __restrict__ Pc,
                                                                const float* __restrict__ bias, const int* __restrict__ nbr_node, int K,
                                                                long long E, float* __restrict__ Y) {
    const long long e = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= E) return;
    const int lane = threadIdx.x & 31;
    const long long i = e / K;
    const int j = __ldg(nbr_node + e);
    float4 z = *reinterpret_cast<const float4*>(Z + e * H + lane * 4);
    const float4 a = *reinterpret_cast<const float4*>(Pa + i * H + lane * 4);
    const float4 c = *reinterpret_cast<const float4*>(Pc + (long long)j * H + lane * 4);
    z.x += a.x + c.x; z.y += a.y + c.y; z.z += a.z + c.z; z.w += a.w + c.w;
    if (bias != nullptr) {
        const float4 b = *reinterpret_cast<const float4*>(bias + lane * 4);
        z.x += b.x; z.y += b.y; z.z += b.z; z.w += b.w;
    }
    *reinterpret_cast<float4*>(Z + e * H + lane * 4) = z;
    *reinterpret_cast<float4*>(Y + e * H + lane * 4) = make_float4(gelu_erf(z.x), gelu_erf(z.y), gelu_erf(z.z), gelu_erf(z.w));
}

// one warp per node: own-half gradient = sum over the node's K edges, gathered-half gradient = sum over the edges pointing at it
__global__ void __launch_bounds__(256) edge_gather_bwd_kernel(const float* __restrict__ dZ, int K, int N, const int* __restrict__ rev_ptr,
                                                              const int* __restrict__ rev_edge, float* __restrict__ dPa, float* __restrict__ dPc) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* own = dZ + (long long)n * K * H + lane * 4;
    for (int k = 0; k < K; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(own + (long long)k * H);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(dPa + (long long)n * H + lane * 4) = s;
    s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = __ldg(rev_ptr + n); t < __ldg(rev_ptr + n + 1); ++t) {
        const float4 v = *reinterpret_cast<const float4*>(dZ + (long long)__ldg(rev_edge + t) * H + lane * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(dPc + (long long)n * H + lane * 4) = s;
}

__global__ void __launch_bounds__(256) masked_sum_fwd_kernel(const float* __restrict__ M, const float* __restrict__ mask_e, int K, int N, float scale,
                                                             float* __restrict__ S) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
        const long long e = (long long)n * K + k;
        const float w = mask_e ? __ldg(mask_e + e) : 1.f;
        const float4 v = *reinterpret_cast<const float4*>(M + e * H + lane * 4);
        s.x += w * v.x; s.y += w * v.y; s.z += w * v.z; s.w += w * v.w;
    }
    *reinterpret_cast<float4*>(S + (long long)n * H + lane * 4) = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
}

__global__ void __launch_bounds__(256) masked_sum_bwd_kernel(const float* __restrict__ dS, const float* __restrict__ mask_e, int K, long long E, float scale,
                                                             float* __restrict__ dM) {
    const long long e = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= E) return;
    const int lane = threadIdx.x & 31;
    const float w = (mask_e ? __ldg(mask_e + e) : 1.f) * scale;
    const float4 v = *reinterpret_cast<const float4*>(dS + (e / K) * H + lane * 4);
    *reinterpret_cast<float4*>(dM + e * H + lane * 4) = make_float4(w * v.x, w * v.y, w * v.z, w * v.w);
}

__global__ void __launch_bounds__(256) row_gather_add_kernel(float* __restrict__ Z, const float* __restrict__ T, const int* __restrict__ idx, long long rows) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    float4 z = *reinterpret_cast<float4*>(Z + r * H + lane * 4);
    const float4 t = *reinterpret_cast<const float4*>(T + (long long)__ldg(idx + r) * H + lane * 4);
    z.x += t.x; z.y += t.y; z.z += t.z; z.w += t.w;
    *reinterpret_cast<float4*>(Z + r * H + lane * 4) = z;
}

// ---- LayerNorm + adaLN modulation ------------------------------------------------------------------------------------------
struct LnModArgs {
    const float *A, *B, *drop;           // x = A + drop * B (B, drop nullable)
    long long rows;
    long long rows_per_member;
    const float *shift, *scale, *gate;   // [members, H] with stride mod_stride (gate nullable)
    long long mod_stride;
    const float* row_mask;               // nullable [rows]
    float eps;
    float *X, *stats, *Y;                // X nullable when B == null (then x = A)
};

__global__ void __launch_bounds__(256) ln_mod_fwd_kernel(const LnModArgs p) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= p.rows) return;
    const int lane = threadIdx.x & 31;
    float4 x4 = *reinterpret_cast<const float4*>(p.A + r * H + lane * 4);
    if (p.B != nullptr) {
        float4 b = *reinterpret_cast<const float4*>(p.B + r * H + lane * 4);
        if (p.drop != nullptr) {
            const float4 d = *reinterpret_cast<const float4*>(p.drop + r * H + lane * 4);
            b.x *= d.x; b.y *= d.y; b.z *= d.z; b.w *= d.w;
        }
        x4.x += b.x; x4.y += b.y; x4.z += b.z; x4.w += b.w;
        if (p.X != nullptr) *reinterpret_cast<float4*>(p.X + r * H + lane * 4) = x4;
    }
    const float v[4] = {x4.x, x4.y, x4.z, x4.w};
    float mean, rstd;
    warp_ln_stats(v, p.eps, mean, rstd);
    if (lane == 0) { p.stats[r * 2] = mean; p.stats[r * 2 + 1] = rstd; }
    const long long m = r / p.rows_per_member;
    const float4 sh = *reinterpret_cast<const float4*>(p.shift + m * p.mod_stride + lane * 4);
    const float4 sc = *reinterpret_cast<const float4*>(p.scale + m * p.mod_stride + lane * 4);
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p.gate != nullptr) g = *reinterpret_cast<const float4*>(p.gate + m * p.mod_stride + lane * 4);
    const float rm = p.row_mask ? __ldg(p.row_mask + r) : 1.f;
    float4 y;
    y.x = rm * g.x * ((v[0] - mean) * rstd * (1.f + sc.x) + sh.x);
    y.y = rm * g.y * ((v[1] - mean) * rstd * (1.f + sc.y) + sh.y);
    y.z = rm * g.z * ((v[2] - mean) * rstd * (1.f + sc.z) + sh.z);
    y.w = rm * g.w * ((v[3] - mean) * rstd * (1.f + sc.w) + sh.w);
    *reinterpret_cast<float4*>(p.Y + r * H + lane * 4) = y;
}

struct LnModBwdArgs {
    const float *dY, *X, *stats;
    long long rows, rows_per_member;
    const float *shift, *scale, *gate;
    long long mod_stride;
    const float* row_mask;
    float* dX;
    float* part;                          // [members][chunks][3][H]: d shift | d scale | d gate partials of one CTA
    int chunks, rows_per_cta;
};

// grid (chunks, members): a CTA owns rows_per_cta consecutive rows of ONE member
__global__ void __launch_bounds__(256) ln_mod_bwd_kernel(const LnModBwdArgs p) {
    __shared__ float sRed[8][3][H];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m = blockIdx.y;
    const long long r_begin = m * p.rows_per_member + (long long)blockIdx.x * p.rows_per_cta;
    const long long r_end = min(m * p.rows_per_member + p.rows_per_member, r_begin + p.rows_per_cta);
    const float4 sh = *reinterpret_cast<const float4*>(p.shift + m * p.mod_stride + lane * 4);
    const float4 sc = *reinterpret_cast<const float4*>(p.scale + m * p.mod_stride + lane * 4);
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p.gate != nullptr) g = *reinterpret_cast<const float4*>(p.gate + m * p.mod_stride + lane * 4);
    const float shv[4] = {sh.x, sh.y, sh.z, sh.w}, scv[4] = {sc.x, sc.y, sc.z, sc.w}, gv[4] = {g.x, g.y, g.z, g.w};
    float d_sh[4] = {0.f, 0.f, 0.f, 0.f}, d_sc[4] = {0.f, 0.f, 0.f, 0.f}, d_g[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long r = r_begin + warp; r < r_end; r += 8) {
        const float4 x4 = *reinterpret_cast<const float4*>(p.X + r * H + lane * 4);
        const float4 dy4 = *reinterpret_cast<const float4*>(p.dY + r * H + lane * 4);
        const float mean = p.stats[r * 2], rstd = p.stats[r * 2 + 1];
        const float rm = p.row_mask ? __ldg(p.row_mask + r) : 1.f;
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, dyv[4] = {dy4.x, dy4.y, dy4.z, dy4.w};
        float xh[4], dxh[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            xh[q] = (xv[q] - mean) * rstd;
            const float gy = rm * dyv[q];
            const float du = gy * gv[q];
            d_g[q] += gy * (xh[q] * (1.f + scv[q]) + shv[q]);
            d_sh[q] += du;
            d_sc[q] += du * xh[q];
            dxh[q] = du * (1.f + scv[q]);
            s1 += dxh[q];
            s2 += dxh[q] * xh[q];
        }
        s1 = warp_sum(s1) * (1.0f / H);
        s2 = warp_sum(s2) * (1.0f / H);
        *reinterpret_cast<float4*>(p.dX + r * H + lane * 4) =
            make_float4(rstd * (dxh[0] - s1 - xh[0] * s2), rstd * (dxh[1] - s1 - xh[1] * s2), rstd * (dxh[2] - s1 - xh[2] * s2), rstd * (dxh[3] - s1 - xh[3] * s2));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { sRed[warp][0][lane * 4 + q] = d_sh[q]; sRed[warp][1][lane * 4 + q] = d_sc[q]; sRed[warp][2][lane * 4 + q] = d_g[q]; }
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * H; t += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sRed[w][t / H][t % H];
        p.part[((m * p.chunks + blockIdx.x) * 3) * H + t] = s;
    }
}

// d shift / d scale / d gate [members, H] (+)= sum over the chunks, fixed order
__global__ void ln_mod_bwd_reduce_kernel(const float* __restrict__ part, int chunks, long long mod_stride, float* d_shift, float* d_scale, float* d_gate,
                                         int accumulate) {
    const int m = blockIdx.x, t = threadIdx.x;            // 3 * H threads
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += part[(((long long)m * chunks + c) * 3) * H + t];
    float* dst = (t < H ? d_shift : (t < 2 * H ? d_scale : d_gate));
    if (dst == nullptr) return;
    dst += (long long)m * mod_stride + (t % H);
    *dst = accumulate ? *dst + s : s;
}

// out[c, :] (+)= sum of the rows whose index is c (embedding gradients: W_s, the positional one-hot).  Two passes, both in a fixed
// order: a CTA walks a contiguous chunk of rows, thread = column, per-class sums in shared memory (a thread owns its column: no
// conflicts, no atomics); then the chunks are added up.
constexpr int IDX_MAX_CLASSES = 72;
__global__ void __launch_bounds__(128) index_sum_part_kernel(const float* __restrict__ X, const int* __restrict__ idx, long long n, int cols, int classes,
                                                             int rows_per_cta, float* __restrict__ part) {
    __shared__ float acc[IDX_MAX_CLASSES][128];
    const int col = threadIdx.x;
    for (int c = 0; c < classes; ++c) acc[c][col] = 0.f;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    if (col < cols)
        for (long long r = r0; r < r1; ++r) acc[__ldg(idx + r)][col] += X[r * cols + col];
    if (col < cols)
        for (int c = 0; c < classes; ++c) part[((long long)blockIdx.x * classes + c) * cols + col] = acc[c][col];
}
__global__ void index_sum_reduce_kernel(const float* __restrict__ part, int n_part, int total, float* __restrict__ out, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float s = 0.f;
    for (int q = 0; q < n_part; ++q) s += part[(long long)q * total + i];
    out[i] = accumulate ? out[i] + s : s;
}

// column sums of X [rows, cols] (bias gradients), two passes, fixed order.  Fast path (cols % 4 == 0, cols <= 1024): thread = one float4
// column group of one row lane, 256 threads cover 256 / (cols / 4) rows per iteration.
__global__ void __launch_bounds__(256) colsum_part_kernel(const float* __restrict__ X, long long rows, int cols, long long ld, int rows_per_cta, float* __restrict__ part) {
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    for (int c = threadIdx.x; c < cols; c += 256) {
        float s = 0.f;
        for (long long r = r0; r < r1; ++r) s += X[r * ld + c];
        part[(long long)blockIdx.x * cols + c] = s;
    }
}
__global__ void __launch_bounds__(256) colsum_part_vec_kernel(const float* __restrict__ X, long long rows, int cols, long long ld, int rows_per_cta, float* __restrict__ part) {
    __shared__ float4 red[256];
    const int groups = cols >> 2;                         // float4 column groups (<= 256)
    const int lanes = 256 / groups;                       // row lanes
    const int cg = threadIdx.x % groups, lane = threadIdx.x / groups;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < lanes)
        for (long long r = r0 + lane; r < r1; r += lanes) {
            const float4 v = *reinterpret_cast<const float4*>(X + r * ld + cg * 4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < groups) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < lanes; ++l) { const float4 v = red[l * groups + threadIdx.x]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }      // fixed order
        *reinterpret_cast<float4*>(part + (long long)blockIdx.x * cols + threadIdx.x * 4) = t;
    }
}
// dpre = dY * GELU'(pre) and the column sums of dpre (the bias gradient of the layer that produced `pre`) in the same pass: the layout of
// colsum_part_vec_kernel (thread = one float4 column group of one row lane), so the sums come out in the same fixed order.
__global__ void __launch_bounds__(256) gelu_bwd_colsum_kernel(const float* __restrict__ pre, const float* __restrict__ dY, long long rows, int cols,
                                                              int rows_per_cta, float* __restrict__ dpre, float* __restrict__ part) {
    __shared__ float4 red[256];
    const int groups = cols >> 2, lanes = 256 / groups;
    const int cg = threadIdx.x % groups, lane = threadIdx.x / groups;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < lanes) {
        // four rows in flight per thread (eight 16-byte loads): the sums are still added in row order
        constexpr int U = 4;
        for (long long r = r0 + lane; r < r1; r += (long long)U * lanes) {
            float4 z[U], g[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long ru = r + (long long)u * lanes;
                if (ru < r1) {
                    z[u] = *reinterpret_cast<const float4*>(pre + ru * cols + cg * 4);
                    g[u] = *reinterpret_cast<const float4*>(dY + ru * cols + cg * 4);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long ru = r + (long long)u * lanes;
                if (ru < r1) {
                    const float4 d = make_float4(g[u].x * gelu_grad(z[u].x), g[u].y * gelu_grad(z[u].y), g[u].z * gelu_grad(z[u].z), g[u].w * gelu_grad(z[u].w));
                    *reinterpret_cast<float4*>(dpre + ru * cols + cg * 4) = d;
                    s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
                }
            }
        }
    }
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < groups) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < lanes; ++l) { const float4 v = red[l * groups + threadIdx.x]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }      // fixed order
        *reinterpret_cast<float4*>(part + (long long)blockIdx.x * cols + threadIdx.x * 4) = t;
    }
}
// second pass of the column sums: block = 32 columns x 8 partial lanes; lane l adds partials l, l + 8, ... in order, the 8 lane sums are then
// added in lane order (fixed order, deterministic)
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ part, int n_part, int cols, float* __restrict__ out, int accumulate) {
    __shared__ float red[8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), l = threadIdx.x >> 5;
    float s = 0.f;
    if (c < cols)
        for (int q = l; q < n_part; q += 8) s += part[(long long)q * cols + c];
    red[l][threadIdx.x & 31] = s;
    __syncthreads();
    if (l == 0 && c < cols) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        out[c] = accumulate ? out[c] + t : t;
    }
}

__global__ void __launch_bounds__(256) sumsq_part_kernel(const float* __restrict__ x, long long n, float* __restrict__ part) {
    __shared__ float sRed[8];
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) s = fmaf(x[i], x[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sRed[w];
        part[blockIdx.x] = t;
    }
}
__global__ void sumsq_final_kernel(const float* __restrict__ part, int n_part, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < n_part; ++q) t += (double)part[q];
        out[0] = (float)t;
    }
}

// torch.optim.AdamW (decoupled weight decay, bias-corrected moments) followed by the EMA of the parameters, one pass over the flat
// buffers.  grad_scale folds the loss scaling / the global-norm clip coefficient; it is read from DEVICE memory (clip_coef[0]) when
// clip_coef != null so that the step needs no host synchronisation: coef = min(1, max_norm / (sqrt(sumsq) + 1e-6)).
__global__ void adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float* __restrict__ ema,
                                 long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, float ema_decay,
                                 const float* __restrict__ sumsq, float max_norm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float coef = 1.f;
    if (sumsq != nullptr && max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(sumsq[0]) + 1e-6f));
    const float gi = g[i] * coef;
    float pi = p[i];
    pi *= 1.f - lr * wd;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
    if (ema != nullptr) ema[i] = ema[i] * ema_decay + pi * (1.f - ema_decay);
}

float* g_work = nullptr;
size_t g_work_bytes = 0;
int workspace(size_t bytes, float** out) {
    if (bytes > g_work_bytes) {
        if (g_work) cudaFree(g_work);
        CB2_CUDA(cudaMalloc(&g_work, bytes));
        g_work_bytes = bytes;
    }
    *out = g_work;
    return 0;
}

}  // namespace
}  // namespace train
}  // namespace cb2

using namespace cb2;
using namespace cb2::train;

extern "C" {

int cb2t_gemm(const float* A, const float* B, float* C, int M, int N, int K, long long lda, long long ldb, long long ldc, int a_kc, int b_kc,
              int accumulate, void* stream) {
    if (!A || !B || !C) { set_error("cb2t_gemm: null argument"); return 1; }
    return gemm(A, B, C, M, N, K, lda, ldb, ldc, a_kc, b_kc, accumulate, (cudaStream_t)stream);
}

int cb2t_set_gemm_mode(int mode) {
    if (mode != 0 && mode != 1) { set_error("cb2t_set_gemm_mode: 0 (fp32 SIMT) or 1 (TF32 tensor cores)"); return 1; }
    g_gemm_mode = mode;
    return 0;
}

int cb2t_bias_gelu_fwd(float* Z, const float* bias, long long rows, int cols, float* Y, void* stream) {
    if (!Z) { set_error("cb2t_bias_gelu_fwd: bad argument"); return 1; }
    const long long n = rows * cols;
    if (cols % 4 != 0) {
        if (Y != nullptr || !bias) { set_error("cb2t_bias_gelu_fwd: the GELU form needs cols % 4 == 0"); return 1; }
        bias_add_scalar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Z, bias, n, cols);
        CB2_LAUNCH_CHECK();
        return 0;
    }
    bias_gelu_fwd_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Z, bias, n, cols, Y);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_linear_bias_gelu_fwd(const float* X, const float* W, const float* bias, float* Z, float* Y, int M, int N, int K, long long ldx, long long ldw,
                              long long ldz, void* stream) {
    if (!X || !W || !bias || !Z) { set_error("cb2t_linear_bias_gelu_fwd: null argument"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    // tensor-core mode: bias and GELU ride in the GEMM's epilogue (Z and Y leave the kernel together)
    if (g_gemm_mode == 1 && N % 32 == 0 && reinterpret_cast<uintptr_t>(bias) % 16 == 0 && (Y == nullptr || reinterpret_cast<uintptr_t>(Y) % 16 == 0) &&
        tc_shape_ok(X, W, Z, M, N, K, ldx, ldw, ldz, 1, 1))
        return gemm_tc_nt_bias_gelu(X, W, bias, Z, Y, M, N, K, ldx, ldw, ldz, s);
    if (ldz != N) { set_error("cb2t_linear_bias_gelu_fwd: the unfused path needs contiguous Z"); return 1; }
    if (int e = gemm(X, W, Z, M, N, K, ldx, ldw, ldz, 1, 1, 0, s)) return e;
    return cb2t_bias_gelu_fwd(Z, bias, M, N, Y, stream);
}

int cb2t_gelu_bwd(const float* pre, const float* dY, long long n, float* dpre, void* stream) {
    if (!pre || !dY || !dpre || n % 4 != 0) { set_error("cb2t_gelu_bwd: bad argument"); return 1; }
    gelu_bwd_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pre, dY, n, dpre);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_elementwise(int mode, const float* a, const float* b, float scale, long long n, float* out, void* stream) {
    if (!a || !out || mode < 0 || mode > 4 || ((mode == 1 || mode == 2 || mode == 4) && !b)) { set_error("cb2t_elementwise: bad argument"); return 1; }
    const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(b);
    if (n % 4 == 0 && al % 16 == 0) {
        elementwise_vec_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mode, a, b, scale, n / 4, out);
        CB2_LAUNCH_CHECK();
        return 0;
    }
    elementwise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mode, a, b, scale, n, out);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_edge_combine_gelu_fwd(float* Z, const float* Pa, const float* Pc, const float* bias, const int* nbr_node, int K, long long E, float* Y, void* stream) {
    if (!Z || !Pa || !Pc || !nbr_node || !Y) { set_error("cb2t_edge_combine_gelu_fwd: null argument"); return 1; }
    edge_combine_gelu_kernel<<<(unsigned)((E + 7) / 8), 256, 0, (cudaStream_t)stream>>>(Z, Pa, Pc, bias, nbr_node, K, E, Y);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_edge_gather_bwd(const float* dZ, int K, int N, const int* rev_ptr, const int* rev_edge, float* dPa, float* dPc, void* stream) {
    if (!dZ || !rev_ptr || !rev_edge || !dPa || !dPc) { set_error("cb2t_edge_gather_bwd: null argument"); return 1; }
    edge_gather_bwd_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(dZ, K, N, rev_ptr, rev_edge, dPa, dPc);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_masked_sum_fwd(const float* M, const float* mask_e, int K, int N, float scale, float* S, void* stream) {
    if (!M || !S) { set_error("cb2t_masked_sum_fwd: null argument"); return 1; }
    masked_sum_fwd_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(M, mask_e, K, N, scale, S);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_masked_sum_bwd(const float* dS, const float* mask_e, int K, long long E, float scale, float* dM, void* stream) {
    if (!dS || !dM) { set_error("cb2t_masked_sum_bwd: null argument"); return 1; }
    masked_sum_bwd_kernel<<<(unsigned)((E + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dS, mask_e, K, E, scale, dM);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_ln_mod_fwd(const float* A, const float* B, const float* drop, long long rows, long long rows_per_member, const float* shift, const float* scale,
                    const float* gate, long long mod_stride, const float* row_mask, float eps, float* X, float* stats, float* Y, void* stream) {
    if (!A || !shift || !scale || !stats || !Y || rows_per_member <= 0 || (B && !X)) { set_error("cb2t_ln_mod_fwd: bad argument"); return 1; }
    LnModArgs p{A, B, drop, rows, rows_per_member, shift, scale, gate, mod_stride, row_mask, eps, X, stats, Y};
    ln_mod_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_ln_mod_bwd(const float* dY, const float* X, const float* stats, long long rows, long long rows_per_member, const float* shift, const float* scale,
                    const float* gate, long long mod_stride, const float* row_mask, float* dX, float* d_shift, float* d_scale, float* d_gate,
                    int accumulate, void* stream) {
    if (!dY || !X || !stats || !shift || !scale || !dX || rows_per_member <= 0 || rows % rows_per_member != 0) { set_error("cb2t_ln_mod_bwd: bad argument"); return 1; }
    const long long members = rows / rows_per_member;
    const int rows_per_cta = 512;
    const int chunks = (int)((rows_per_member + rows_per_cta - 1) / rows_per_cta);
    float* part = nullptr;
    if (int e = workspace((size_t)members * chunks * 3 * H * sizeof(float), &part)) return e;
    LnModBwdArgs p{dY, X, stats, rows, rows_per_member, shift, scale, gate, mod_stride, row_mask, dX, part, chunks, rows_per_cta};
    ln_mod_bwd_kernel<<<dim3(chunks, (unsigned)members), 256, 0, (cudaStream_t)stream>>>(p);
    CB2_LAUNCH_CHECK();
    ln_mod_bwd_reduce_kernel<<<(unsigned)members, 3 * H, 0, (cudaStream_t)stream>>>(part, chunks, mod_stride, d_shift, d_scale, gate ? d_gate : nullptr, accumulate);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_edge_raw_features(const float* X, const int* idx, const float* D, int F, int L, int K, float* raw, void* stream) {
    if (!X || !idx || !D || !raw) { set_error("cb2t_edge_raw_features: null argument"); return 1; }
    return launch_edge_raw_features(X, idx, D, F, L, K, raw, (cudaStream_t)stream);
}

int cb2t_row_gather_add(float* Z, const float* T, const int* idx, long long rows, void* stream) {
    if (!Z || !T || !idx) { set_error("cb2t_row_gather_add: null argument"); return 1; }
    row_gather_add_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(Z, T, idx, rows);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_index_sum(const float* X, const int* idx, long long n, int cols, int classes, float* out, int accumulate, void* stream) {
    if (!X || !idx || !out || cols <= 0 || cols > 128 || classes <= 0 || classes > IDX_MAX_CLASSES) {
        set_error("cb2t_index_sum: cols <= 128 and classes <= %d", IDX_MAX_CLASSES);
        return 1;
    }
    const int n_part = (int)(n < 592 * 64 ? (n + 63) / 64 : 592);
    const int rows_per_cta = (int)((n + n_part - 1) / n_part);
    float* part = nullptr;
    if (int e = workspace((size_t)n_part * classes * cols * sizeof(float), &part)) return e;
    index_sum_part_kernel<<<n_part, 128, 0, (cudaStream_t)stream>>>(X, idx, n, cols, classes, rows_per_cta, part);
    CB2_LAUNCH_CHECK();
    const int total = classes * cols;
    index_sum_reduce_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(part, n_part, total, out, accumulate);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_colsum(const float* X, long long rows, int cols, long long ld, float* out, int accumulate, void* stream) {
    if (!X || !out) { set_error("cb2t_colsum: null argument"); return 1; }
    const int n_part = (int)(rows < 1184 * 32 ? (rows + 31) / 32 : 1184);
    const int rows_per_cta = (int)((rows + n_part - 1) / n_part);
    float* part = nullptr;
    if (int e = workspace((size_t)n_part * cols * sizeof(float), &part)) return e;
    const bool vec = cols % 4 == 0 && cols <= 1024 && 256 % (cols / 4) == 0 && ld % 4 == 0 && reinterpret_cast<uintptr_t>(X) % 16 == 0;
    if (vec) colsum_part_vec_kernel<<<n_part, 256, 0, (cudaStream_t)stream>>>(X, rows, cols, ld, rows_per_cta, part);
    else colsum_part_kernel<<<n_part, 256, 0, (cudaStream_t)stream>>>(X, rows, cols, ld, rows_per_cta, part);
    CB2_LAUNCH_CHECK();
    colsum_reduce_kernel<<<(cols + 31) / 32, 256, 0, (cudaStream_t)stream>>>(part, n_part, cols, out, accumulate);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_gelu_bwd_colsum(const float* pre, const float* dY, long long rows, int cols, float* dpre, float* colsum_out, int accumulate, void* stream) {
    if (!pre || !dY || !dpre || !colsum_out) { set_error("cb2t_gelu_bwd_colsum: null argument"); return 1; }
    if (cols % 4 != 0 || cols > 1024 || 256 % (cols / 4) != 0) { set_error("cb2t_gelu_bwd_colsum: cols=%d unsupported", cols); return 1; }
    const int n_part = (int)(rows < 1184 * 32 ? (rows + 31) / 32 : 1184);
    const int rows_per_cta = (int)((rows + n_part - 1) / n_part);
    float* part = nullptr;
    if (int e = workspace((size_t)n_part * cols * sizeof(float), &part)) return e;
    gelu_bwd_colsum_kernel<<<n_part, 256, 0, (cudaStream_t)stream>>>(pre, dY, rows, cols, rows_per_cta, dpre, part);
    CB2_LAUNCH_CHECK();
    colsum_reduce_kernel<<<(cols + 31) / 32, 256, 0, (cudaStream_t)stream>>>(part, n_part, cols, colsum_out, accumulate);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_sumsq(const float* x, long long n, float* out, void* stream) {
    if (!x || !out) { set_error("cb2t_sumsq: null argument"); return 1; }
    const int n_part = 592;
    float* part = nullptr;
    if (int e = workspace((size_t)n_part * sizeof(float), &part)) return e;
    sumsq_part_kernel<<<n_part, 256, 0, (cudaStream_t)stream>>>(x, n, part);
    CB2_LAUNCH_CHECK();
    sumsq_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(part, n_part, out);
    CB2_LAUNCH_CHECK();
    return 0;
}

int cb2t_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, long long n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, float ema_decay, const float* grad_sumsq, float max_norm, void* stream) {
    if (!p || !g || !m || !v || step < 1) { set_error("cb2t_adamw_ema: bad argument"); return 1; }
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adamw_ema_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, ema_decay,
                                                                                   grad_sumsq, max_norm);
    CB2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
