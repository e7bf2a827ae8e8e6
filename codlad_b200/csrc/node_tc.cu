// tcgen05 tier of the per-node kernels (same maths as node.cu; reference models/protein_mpnn_utils.py:247-259,
// :307-317, models/latent_model.py:21-35,214, diffusion_and_flow/gaussian_diffusion.py:303-318,345-351,440-446).
//
// One CTA = one tile of 128 nodes.  Warps 0-15 are the epilogue warps: thread (row r, column quarter cq) owns 32 of the
// 128 columns of its node row (the four warps that can address a TMEM lane quarter split the columns); LayerNorm
// row statistics are exchanged between the four quarters through drained TMEM accumulator columns.
// Warp 16 is the control warp: one lane streams the layer's fp16 weight blocks through four 32 KB shared-memory
// slots with TMA and issues the tcgen05 MMAs.  The node kernels are latency chains (W3 -> LN -> FFN -> LN ->
// projections), so the CTA alternates strictly between an MMA phase and an epilogue phase:
//
//   P1  acc0 = S . W3^T                      EA  h1 = gate1 * mod(LN(h_V + (acc0 + cnt b3)/30))
//   P2  acc1|acc2 = h1 . Win[0:256]^T        EG  mid = GELU(acc + b_in)
//   P3  acc3 = mid . Wout[:, 0:256]^T ; acc1|acc2 = h1 . Win[256:512]^T      EG
//   P4  acc3 += mid . Wout[:, 256:512]^T     EB  h2 = mask * gate2 * mod(LN(h1 + acc3 + b_out))   (+ FinalLayer, p_sample)
//   P5  own / gathered halves of the next edge MLPs' first layer                EP  -> P (fp32 own half), Pc (fp16 gathered half)
//
// State h_V stays fp32 in HBM; only MMA operands are fp16 (fp32 accumulation in TMEM).
#include "model.h"
#include "tc_common.cuh"

namespace cb2 {

using namespace tc;

namespace {

struct ProjTc {
    int wa_row, wc_row;          // weight blocks (rows of the packed fp16 weight tensor)
    const float *ba, *table;     // own-half bias; optional per-residue-type table added to the gathered half
    __half* out16;               // [N, 256] fp16: [own half + bias | gathered half (+ table)]
    int add_enc;                 // 0: h' = h, 1: h' = h + hVenc, 2: h' = 2 h
};

struct NodeTcParams {
    int N, L, K;
    int do_update, masked_count;
    const float *x, *xin_w_t, *xin_b;
    const float* S;
    int w3_row, win_row, wout_row;           // win/wout: 4 consecutive 128-row blocks each
    const float *b3, *bin, *bout;
    const float* mod;
    int mod_stride;
    const int *lengths, *frame_of, *nbr_idx, *cg_z;
    float *hV, *hVenc;
    int write_enc;
    ProjTc proj[2];
    int n_proj;
    int do_final;
    const float *fin_mod, *fin_w_t, *fin_b;
    float* out6;
    const float *x_t, *noise, *coef;
    float* x_next;
    unsigned long long* trace;   // debug timeline of CTA 0 (nullptr = off)
    int trace_slot;              // debug: launch window slot (trace_window)
};

__device__ __forceinline__ void ldg_f32x8(const float* p, float* v) {
    uint32_t r[8];
    ldg256_coherent(p, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void stg_f32x8(float* p, const float* v) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(v[i]);
    stg256(p, r);
}
__device__ __forceinline__ void ld32(const float* p, float* v) {       // 32 consecutive floats, same address in every lane
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
        v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
    }
}
// write 32 consecutive columns [c0, c0+32) of row r as fp16 into a swizzled K-major operand tile
__device__ __forceinline__ void store_row_chunk(unsigned char* tile, int r, int c0, const float* v) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = f2_to_h2(v[u * 8 + e * 2], v[u * 8 + e * 2 + 1]);
        *reinterpret_cast<uint4*>(tile + tile_off(r, (c0 >> 3) + u)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

constexpr int NODE_EPI_THREADS = 512;

__device__ __forceinline__ void node_epi_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__global__ void __launch_bounds__(NODE_EPI_THREADS + 32, 1) node_tc_kernel(const __grid_constant__ CUtensorMap wmap, const NodeTcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* hA = smem;                         // current node state as the A operand
    unsigned char* mid0 = hA + TILE_BYTES;            // FFN hidden block / h' of the decoder projections
    unsigned char* mid1 = mid0 + TILE_BYTES;
    unsigned char* sW = mid1 + TILE_BYTES;            // 4 weight slots
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sW + 4 * TILE_BYTES);      // [0] weights full, [1] mma done, [2] activations ready
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) trace_window(p.trace, p.trace_slot, false);
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    pdl_launch_dependents();               // the next kernel of the step may start its prologue on the SMs this grid leaves idle
    const uint32_t bar_full = smem_u32(&sBar[0]), bar_mma = smem_u32(&sBar[1]), bar_act = smem_u32(&sBar[2]);
    if (tid == 0) {
        mbar_init(bar_full, 1); mbar_init(bar_mma, 1); mbar_init(bar_act, NODE_EPI_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;
    constexpr uint32_t IDESC = umma_idesc(128, 128, 0, 0);
    const int n_phases = (p.do_update ? 4 : 0) + (p.n_proj > 0 ? 1 : 0);

    if (warp == 16) {
        // ------------------------------------------------------------------ control warp
        if (lane == 0) {
            unsigned long long* ctrace = (p.trace != nullptr && blockIdx.x == 0) ? p.trace + 512 : nullptr;
            int n_ct = 0;
            auto cmark = [&](int ev) {
                if (ctrace != nullptr && n_ct < 400) {
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    ctrace[1 + n_ct] = (t << 8) | (unsigned long long)ev;
                    ctrace[0] = (unsigned long long)(++n_ct);
                }
            };
            cmark(0);
            uint32_t ph_full = 0, ph_mma = 0, ph_act = 0;
            auto load_weights = [&](const int* rows, int n) {
                mbar_expect_tx(bar_full, (uint32_t)(n * TILE_BYTES));
                for (int i = 0; i < n; ++i)
                    for (int h = 0; h < 2; ++h)
                        tma_load_2d(smem_u32(sW + i * TILE_BYTES + h * HALF_BYTES), &wmap, h * 64, rows[i], bar_full);
            };
            auto mma = [&](const unsigned char* a, int slot, int region, bool accumulate) {
                const uint32_t a_u = smem_u32(a), b_u = smem_u32(sW + slot * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t koff = (uint32_t)((k >> 2) * HALF_BYTES + (k & 3) * 32);
                    umma_f16(tmem_base + (uint32_t)(region * 128), umma_desc(a_u + koff, 16, 1024), umma_desc(b_u + koff, 16, 1024), IDESC,
                             (accumulate || k > 0) ? 1u : 0u);
                }
            };
            for (int ph = 0; ph < n_phases; ++ph) {
                const int kind = p.do_update ? ph : 4;           // 0..3 = P1..P4, 4 = projections
                int rows[4], n = 0;
                if (kind == 0) { rows[n++] = p.w3_row; }
                else if (kind == 1) { rows[n++] = p.win_row; rows[n++] = p.win_row + 128; }
                else if (kind == 2) { rows[n++] = p.wout_row; rows[n++] = p.wout_row + 128; rows[n++] = p.win_row + 256; rows[n++] = p.win_row + 384; }
                else if (kind == 3) { rows[n++] = p.wout_row + 256; rows[n++] = p.wout_row + 384; }
                else { for (int j = 0; j < p.n_proj; ++j) { rows[n++] = p.proj[j].wa_row; rows[n++] = p.proj[j].wc_row; } }
                load_weights(rows, n);                           // slots are free: the previous phase's MMAs have completed
                cmark(1);
                mbar_wait(bar_act, ph_act); ph_act ^= 1;         // operands written, accumulators drained
                cmark(2);
                mbar_wait(bar_full, ph_full); ph_full ^= 1;
                cmark(3);
                tc_fence_after();
                if (kind == 0) { mma(hA, 0, 0, false); }
                else if (kind == 1) { mma(hA, 0, 1, false); mma(hA, 1, 2, false); }
                else if (kind == 2) { mma(mid0, 0, 3, false); mma(mid1, 1, 3, true); mma(hA, 2, 1, false); mma(hA, 3, 2, false); }
                else if (kind == 3) { mma(mid0, 0, 3, true); mma(mid1, 1, 3, true); }
                else {
                    for (int j = 0; j < p.n_proj; ++j) {
                        mma(hA, 2 * j, 2 * j, false);
                        mma(p.proj[j].add_enc ? mid0 : hA, 2 * j + 1, 2 * j + 1, false);
                    }
                }
                umma_commit(bar_mma);
                cmark(4);
                mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;         // weight slots and operand tiles may be overwritten
                cmark(5);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps: thread = (node row, 32 columns)
        const int quarter = warp & 3, cq = warp >> 2, r = quarter * 32 + lane, c0 = cq * 32;
        const int n = blockIdx.x * 128 + r;
        const bool live = n < p.N;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
        uint32_t ph_mma = 0;
        int b = 0, zres = 0;
        float mk = 0.f, cnt = (float)p.K;
        // per-row metadata (member, residue type, node mask, masked-neighbour count): dependent global loads, deliberately
        // issued AFTER the first operand stage so they do not delay the first MMA
        auto load_row_meta = [&]() {
            if (!live) return;
            b = n / p.L;
            const int i = n - b * p.L;
            const int f = __ldg(p.frame_of + b);
            const int len = __ldg(p.lengths + f);
            mk = i < len ? 1.f : 0.f;
            zres = __ldg(p.cg_z + (size_t)f * p.L + i);
            if (p.masked_count && len < p.L) {
                int cn = 0;
                if (i < len) {
                    const int* row = p.nbr_idx + ((size_t)f * p.L + i) * p.K;
                    for (int k = 0; k < p.K; ++k) cn += __ldg(row + k) < len ? 1 : 0;
                }
                cnt = (float)cn;
            }
        };
        const float* m = p.mod + (size_t)(live ? n / p.L : 0) * p.mod_stride + c0;
        float v[32];                                    // this thread's 32 columns of the row state, fp32
        pdl_wait();                                     // S / h_V / x are produced by the previous kernels of the step
        auto publish = [&]() { fence_async_smem(); tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(bar_act); };    // one arrival per warp
        auto wait_mma = [&]() { mbar_wait(bar_mma, ph_mma); ph_mma ^= 1; tc_fence_after(); };
        auto store16 = [&](unsigned char* tile, int g16, const float* x) {     // 16 columns [c0 + 16 g16, +16) of row r as fp16
            uint32_t o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = f2_to_h2(x[e * 2], x[e * 2 + 1]);
            *reinterpret_cast<uint4*>(tile + tile_off(r, (c0 >> 3) + g16 * 2)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(tile + tile_off(r, (c0 >> 3) + g16 * 2 + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
        };
        // LayerNorm statistics of the full row (no affine, eps 1e-6): the four column quarters of a row exchange their partial
        // sums through accumulator columns this thread has just drained (`region`): tcgen05.st, one barrier, tcgen05.ld; the
        // four partials are summed in a fixed order, so the result is deterministic
        auto row_stats = [&](int region, float& mean, float& rstd) {
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) { sum += v[c]; sq = fmaf(v[c], v[c], sq); }
            tmem_st2(tmem_lane + (uint32_t)(region * 128), sum, sq);
            tc_fence_before();
            node_epi_sync();
            tc_fence_after();
            float part[8];
            tmem_ld2_x4(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(region * 128), 32u, part);
            const float tsum = (part[0] + part[2]) + (part[4] + part[6]), tsq = (part[1] + part[3]) + (part[5] + part[7]);
            mean = tsum * (1.0f / 128.0f);
            rstd = rsqrtf(fmaxf(tsq * (1.0f / 128.0f) - mean * mean, 0.f) + 1e-6f);
        };

        if (p.do_update) {
            // ---- E0: S -> fp16 A operand; h_V -> v ----
            {
                float s32[32];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (live) { ldg_f32x8(p.S + (size_t)n * 128 + c0 + u * 8, s32 + u * 8); ldg_f32x8(p.hV + (size_t)n * 128 + c0 + u * 8, v + u * 8); }
                    else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) { s32[u * 8 + e] = 0.f; v[u * 8 + e] = 0.f; }
                    }
                }
                store16(hA, 0, s32); store16(hA, 1, s32 + 16);
            }
            publish();
            load_row_meta();
            // ---- EA: h1 = gate1 * (LN(h_V + (acc + cnt b3)/30) (1 + scale1) + shift1) ----
            wait_mma();
#pragma unroll
            for (int g16 = 0; g16 < 2; ++g16) {
                float acc[16], bb[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) { const float4 t = __ldg(reinterpret_cast<const float4*>(p.b3 + c0 + g16 * 16) + q); bb[q * 4] = t.x; bb[q * 4 + 1] = t.y; bb[q * 4 + 2] = t.z; bb[q * 4 + 3] = t.w; }
                tmem_ld16(tmem_lane + (uint32_t)(g16 * 16), acc);
#pragma unroll
                for (int e = 0; e < 16; ++e) v[g16 * 16 + e] += (acc[e] + cnt * bb[e]) * (1.0f / 30.0f);
            }
            {
                float mean, rstd;
                row_stats(0, mean, rstd);
                float sh[32], sc[32], gt[32];
                ld32(m, sh); ld32(m + 128, sc); ld32(m + 256, gt);
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = gt[c] * fmaf((v[c] - mean) * rstd, 1.0f + sc[c], sh[c]);
                store16(hA, 0, v); store16(hA, 1, v + 16);
            }
            publish();
            // ---- EG x2: FFN hidden = GELU(h1 Win^T + b_in), 256 columns per phase ----
            for (int half = 0; half < 2; ++half) {
                wait_mma();
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    unsigned char* dst = blk ? mid1 : mid0;
                    const float* bi = p.bin + half * 256 + blk * 128 + c0;
#pragma unroll
                    for (int g16 = 0; g16 < 2; ++g16) {
                        float acc[16], bb[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) { const float4 t = __ldg(reinterpret_cast<const float4*>(bi + g16 * 16) + q); bb[q * 4] = t.x; bb[q * 4 + 1] = t.y; bb[q * 4 + 2] = t.z; bb[q * 4 + 3] = t.w; }
                        tmem_ld16(tmem_lane + (uint32_t)((1 + blk) * 128 + g16 * 16), acc);
                        uint32_t o[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = as_u32(gelu2_h2(as_h2(pack_sat(acc[e * 2] + bb[e * 2], acc[e * 2 + 1] + bb[e * 2 + 1]))));
                        *reinterpret_cast<uint4*>(dst + tile_off(r, (c0 >> 3) + g16 * 2)) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4*>(dst + tile_off(r, (c0 >> 3) + g16 * 2 + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
                publish();
            }
            // ---- EB: h2 = mask * gate2 * (LN(h1 + acc3 + b_out) (1 + scale2) + shift2) ----
            wait_mma();
#pragma unroll
            for (int g16 = 0; g16 < 2; ++g16) {
                float acc[16], bb[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) { const float4 t = __ldg(reinterpret_cast<const float4*>(p.bout + c0 + g16 * 16) + q); bb[q * 4] = t.x; bb[q * 4 + 1] = t.y; bb[q * 4 + 2] = t.z; bb[q * 4 + 3] = t.w; }
                tmem_ld16(tmem_lane + (uint32_t)(3 * 128 + g16 * 16), acc);
#pragma unroll
                for (int e = 0; e < 16; ++e) v[g16 * 16 + e] += acc[e] + bb[e];
            }
            {
                float mean, rstd;
                row_stats(3, mean, rstd);
                float sh[32], sc[32], gt[32];
                ld32(m + 384, sh); ld32(m + 512, sc); ld32(m + 640, gt);
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = mk * (gt[c] * fmaf((v[c] - mean) * rstd, 1.0f + sc[c], sh[c]));
            }
        } else {
            load_row_meta();
            // ---- node init: h = x_in(x) ----
            float x0 = 0.f, x1 = 0.f, x2 = 0.f;
            if (live) { x0 = p.x[(size_t)n * 3]; x1 = p.x[(size_t)n * 3 + 1]; x2 = p.x[(size_t)n * 3 + 2]; }
            float w0[32], w1[32], w2[32], bb[32];
            ld32(p.xin_w_t + c0, w0); ld32(p.xin_w_t + 128 + c0, w1); ld32(p.xin_w_t + 256 + c0, w2); ld32(p.xin_b + c0, bb);
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = live ? fmaf(x2, w2[e], fmaf(x1, w1[e], fmaf(x0, w0[e], bb[e]))) : 0.f;
        }
        // ---- write h_V (and the decoder's frozen encoder state), stage the projection operands ----
        const int enc_mode = p.n_proj > 0 ? p.proj[p.n_proj - 1].add_enc : 0;     // only the decoder-facing projection uses h'
        if (live) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                stg_f32x8(p.hV + (size_t)n * 128 + c0 + u * 8, v + u * 8);
                if (p.write_enc) stg_f32x8(p.hVenc + (size_t)n * 128 + c0 + u * 8, v + u * 8);
            }
        }
        if (p.n_proj > 0) {
            store16(hA, 0, v); store16(hA, 1, v + 16);
            if (enc_mode) {
                float hp[32];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float e8[8];
                    if (enc_mode == 1 && live) ldg_f32x8(p.hVenc + (size_t)n * 128 + c0 + u * 8, e8);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float hv = v[u * 8 + e];
                        hp[u * 8 + e] = enc_mode == 1 ? (live ? hv + e8[e] : 0.f) : 2.0f * hv;
                    }
                }
                store16(mid0, 0, hp); store16(mid0, 1, hp + 16);
            }
            publish();
            // ---- EP: own halves (+ bias) and gathered halves (+ residue-type table) -> P16 (fp16) ----
            wait_mma();
            for (int j = 0; j < p.n_proj; ++j) {
                const ProjTc pj = p.proj[j];
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    float acc[32], add[32];
                    if (which == 0) ld32(pj.ba + c0, add);
                    else if (pj.table != nullptr) ld32(pj.table + zres * 128 + c0, add);
                    else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) add[e] = 0.f;
                    }
                    tmem_ld32(tmem_lane + (uint32_t)((2 * j + which) * 128), acc);
                    if (live) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            uint32_t o[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) o[e] = f2_to_h2(acc[u * 16 + e * 2] + add[u * 16 + e * 2], acc[u * 16 + e * 2 + 1] + add[u * 16 + e * 2 + 1]);
                            stg256(pj.out16 + (size_t)n * 256 + which * 128 + c0 + u * 16, o);
                        }
                    }
                }
            }
        }
        if (p.do_final) {
            // ---- FinalLayer (adaLN modulate + Linear 128 -> 6) fused with the DDPM p_sample update: one thread per row
            //      re-reads the full row it and its three column-quarter partners just wrote ----
            node_epi_sync();
            if (cq == 0 && live) {
                const float* hrow = p.hV + (size_t)n * 128;
                float sum = 0.f, sq = 0.f;
#pragma unroll 1
                for (int u = 0; u < 16; ++u) {
                    float h8[8];
                    ldg_f32x8(hrow + u * 8, h8);
#pragma unroll
                    for (int e = 0; e < 8; ++e) { sum += h8[e]; sq = fmaf(h8[e], h8[e], sq); }
                }
                const float mean = sum * (1.0f / 128.0f);
                const float rstd = rsqrtf(fmaxf(sq * (1.0f / 128.0f) - mean * mean, 0.f) + 1e-6f);
                const float* fm = p.fin_mod + (size_t)b * p.mod_stride;      // [shift | scale]
                float o[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
                for (int u = 0; u < 16; ++u) {
                    float h8[8];
                    ldg_f32x8(hrow + u * 8, h8);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int c = u * 8 + e;
                        const float t = fmaf((h8[e] - mean) * rstd, 1.0f + __ldg(fm + 128 + c), __ldg(fm + c));
                        const float2* w = reinterpret_cast<const float2*>(p.fin_w_t + c * 6);
                        const float2 w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
                        o[0] = fmaf(t, w0.x, o[0]); o[1] = fmaf(t, w0.y, o[1]); o[2] = fmaf(t, w1.x, o[2]);
                        o[3] = fmaf(t, w1.y, o[3]); o[4] = fmaf(t, w2.x, o[4]); o[5] = fmaf(t, w2.y, o[5]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 6; ++u) { o[u] += __ldg(p.fin_b + u); p.out6[(size_t)n * 6 + u] = o[u]; }
                if (p.x_next != nullptr) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float eps = o[d], vv = o[3 + d];
                        const float x = p.x_t[(size_t)n * 3 + d];
                        const float frac = (vv + 1.0f) / 2.0f;
                        const float logvar = frac * p.coef[1] + (1.0f - frac) * p.coef[0];
                        const float x0 = p.coef[2] * x - p.coef[3] * eps;
                        const float mean_ = p.coef[4] * x0 + p.coef[5] * x;
                        p.x_next[(size_t)n * 3 + d] = mean_ + p.coef[6] * expf(0.5f * logvar) * p.noise[(size_t)n * 3 + d];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) trace_window(p.trace, p.trace_slot, true);
    if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

constexpr size_t NODE_TC_SMEM = 7 * (size_t)TILE_BYTES + 64;

struct NodeTcState { CUtensorMap wmap; };

int node_tc_launch(Plan& p, NodeTcParams& np, cudaStream_t s) {
    const NodeTcState& st = *reinterpret_cast<const NodeTcState*>(p.node_tc);
    CB2_CUDA(launch_pdl(node_tc_kernel, dim3((np.N + 127) / 128), dim3(NODE_EPI_THREADS + 32), NODE_TC_SMEM, s, st.wmap, np));
    CB2_LAUNCH_CHECK();
    p.launches++;
    return 0;
}

void base_params(Plan& p, NodeTcParams& np) {
    np = NodeTcParams{};
    np.N = p.NB * p.L; np.L = p.L; np.K = p.K;
    np.lengths = p.lengths; np.frame_of = p.frame_of; np.nbr_idx = p.nbr_idx; np.cg_z = p.cg_z;
    np.hV = p.hV; np.hVenc = p.hVenc; np.S = p.S;
    np.trace = p.tc_trace; np.trace_slot = p.launches & 2047;
}

}  // namespace

int node_tc_prepare(Plan& p) {
    EncodeTiledFn fn = nullptr;
    if (get_encode_fn(&fn)) return 1;
    NodeTcState* st = new NodeTcState();
    p.node_tc = st;
    if (encode_rows_map(fn, &st->wmap, p.model->dev_f16, (size_t)p.model->n_f16_blocks * 128, 128)) return 1;
    CB2_CUDA(cudaFuncSetAttribute(node_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NODE_TC_SMEM));
    return 0;
}

void node_tc_release(Plan& p) {
    delete reinterpret_cast<NodeTcState*>(p.node_tc);
    p.node_tc = nullptr;
}

int launch_node_init_tc(Plan& p, const float* x, cudaStream_t s) {
    const DenoiserModel& m = *p.model;
    auto row_of = [&](const __half* w) { return (int)((w - m.dev_f16) / 128); };
    NodeTcParams np;
    base_params(p, np);
    np.do_update = 0;
    np.x = x; np.xin_w_t = m.xin_w_t; np.xin_b = m.xin_b;
    np.n_proj = 1;
    np.proj[0] = ProjTc{row_of(m.enc[0].W1a_h), row_of(m.enc[0].W1c_h), m.enc[0].b1, nullptr, p.P16[0], 0};
    return node_tc_launch(p, np, s);
}

int launch_node_update_tc(Plan& p, int phase, const float* mod_base, int mod_stride_b, const float* x_t, const float* noise,
                          float* x_next, const float* coef_row, cudaStream_t s) {
    const DenoiserModel& m = *p.model;
    auto row_of = [&](const __half* w) { return (int)((w - m.dev_f16) / 128); };
    NodeTcParams np;
    base_params(p, np);
    np.do_update = 1;
    np.mod_stride = mod_stride_b;
    if (phase < 3) {
        const EncLayerW& e = m.enc[phase];
        np.masked_count = 1;
        np.w3_row = row_of(e.W3_h); np.win_row = row_of(e.Win_h); np.wout_row = row_of(e.Wout_h);
        np.b3 = e.b3; np.bin = e.bin; np.bout = e.bout;
        np.mod = mod_base + CB2_MOD_ENC_OFF(phase);
        np.n_proj = 2;
        np.proj[0] = ProjTc{row_of(e.W11a_h), row_of(e.W11c_h), e.b11, nullptr, p.P16[1], 0};
        if (phase < 2) {
            const EncLayerW& nx = m.enc[phase + 1];
            np.proj[1] = ProjTc{row_of(nx.W1a_h), row_of(nx.W1c_h), nx.b1, nullptr, p.P16[0], 0};
        } else {
            const DecLayerW& d = m.dec[0];
            np.write_enc = 1;
            np.proj[1] = ProjTc{row_of(d.W1a_h), row_of(d.W1d_h), d.b1, d.TS, p.P16[0], 2};
        }
    } else {
        const int l = phase - 3;
        const DecLayerW& d = m.dec[l];
        np.masked_count = 0;
        np.w3_row = row_of(d.W3_h); np.win_row = row_of(d.Win_h); np.wout_row = row_of(d.Wout_h);
        np.b3 = d.b3; np.bin = d.bin; np.bout = d.bout;
        np.mod = mod_base + CB2_MOD_DEC_OFF(l);
        if (l < 2) {
            const DecLayerW& nx = m.dec[l + 1];
            np.n_proj = 1;
            np.proj[0] = ProjTc{row_of(nx.W1a_h), row_of(nx.W1d_h), nx.b1, nx.TS, p.P16[0], 1};
        } else {
            np.n_proj = 0;
            np.do_final = 1;
            np.fin_mod = mod_base + CB2_MOD_FIN_OFF;
            np.fin_w_t = m.fin_w_t; np.fin_b = m.fin_b;
            np.out6 = p.out6;
            np.x_t = x_t; np.noise = noise; np.x_next = x_next; np.coef = coef_row;
        }
    }
    return node_tc_launch(p, np, s);
}

}  // namespace cb2
