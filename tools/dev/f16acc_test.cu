// Probe: layout of an fp16 accumulator (kind::f16, D = f16) in tensor memory, as seen by tcgen05.ld 32x32b.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../codlad_b200/csrc/tc_common.cuh"
using namespace cb2::tc;
namespace cb2 { void set_error(const char*, ...) {} }

__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, uint32_t* out, uint32_t* out_pack, int f16_acc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem; unsigned char* sB = smem + TILE_BYTES;
    __shared__ uint64_t bar; __shared__ uint32_t tm;
    const int tid = threadIdx.x;
    for (int i = tid; i < 128 * 16; i += 128) {
        const int r = i >> 4, c16 = i & 15;
        *reinterpret_cast<uint4*>(sA + tile_off(r, c16)) = *reinterpret_cast<const uint4*>(A + r * 128 + c16 * 8);
        *reinterpret_cast<uint4*>(sB + tile_off(r, c16)) = *reinterpret_cast<const uint4*>(B + r * 128 + c16 * 8);
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tm;
    if (tid < 32) {
        const uint32_t idesc = (f16_acc ? 0u : (1u << 4)) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (elect_one()) {
            for (int k = 0; k < 8; ++k) {
                const uint32_t koff = (uint32_t)((k >> 2) * HALF_BYTES + (k & 3) * 32);
                umma_f16(tmem, umma_desc(smem_u32(sA) + koff, 16, 1024), umma_desc(smem_u32(sB) + koff, 16, 1024), idesc, k > 0);
            }
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
    for (int c = 0; c < 128; c += 16) {
        uint32_t r[16];

        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(lane_addr + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 16; ++e) out[tid * 128 + c + e] = r[e];
    }
    // packed load: 16 registers <- 32 columns of 16-bit data
    for (int c = 0; c < 128; c += 32) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.pack::16b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(lane_addr + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 16; ++e) out_pack[tid * 64 + c / 2 + e] = r[e];
    }
    tc_fence_before(); __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<__half> hA(128 * 128), hB(128 * 128);
    std::vector<float> fA(128 * 128), fB(128 * 128);
    srand(1);
    for (int i = 0; i < 128 * 128; ++i) {
        fA[i] = (rand() % 2001 - 1000) / 1000.0f; fB[i] = (rand() % 2001 - 1000) / 4000.0f;
        hA[i] = __float2half(fA[i]); hB[i] = __float2half(fB[i]); fA[i] = __half2float(hA[i]); fB[i] = __half2float(hB[i]);
    }
    __half *dA, *dB; uint32_t *dO, *dP;
    cudaMalloc(&dA, 32768); cudaMalloc(&dB, 32768); cudaMalloc(&dO, 128 * 128 * 4); cudaMalloc(&dP, 128 * 64 * 4);
    cudaMemcpy(dA, hA.data(), 32768, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), 32768, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES + 1024);
    std::vector<double> ref(128 * 128);
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 128; ++c) { double s = 0; for (int k = 0; k < 128; ++k) s += (double)fA[r * 128 + k] * fB[c * 128 + k]; ref[r * 128 + c] = s; }
    for (int f16 = 0; f16 < 2; ++f16) {
        cudaMemset(dO, 0, 128 * 128 * 4); cudaMemset(dP, 0, 128 * 64 * 4);
        probe<<<1, 128, 2 * TILE_BYTES + 1024>>>(dA, dB, dO, dP, f16);
        cudaError_t e = cudaDeviceSynchronize();
        printf("== accumulate in %s: %s\n", f16 ? "f16" : "f32", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint32_t> o(128 * 128), pk(128 * 64);
        cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(pk.data(), dP, pk.size() * 4, cudaMemcpyDeviceToHost);
        printf("row 5 raw cells 0..7: "); for (int c = 0; c < 8; ++c) printf("%08x ", o[5 * 128 + c]); printf("\n");
        printf("row 5 ref      0..7: "); for (int c = 0; c < 8; ++c) printf("%8.4f ", ref[5 * 128 + c]); printf("\n");
        if (!f16) { double me = 0; for (int i = 0; i < 128 * 128; ++i) { float v; memcpy(&v, &o[i], 4); me = fmax(me, fabs(v - ref[i])); } printf("f32 max abs err %.3g\n", me); }
        else {
            // hypothesis A: cell c holds element c in its low half; hypothesis B: cell c holds elements (2c, 2c+1)
            double ea = 0, eb = 0, ep = 0;
            for (int r = 0; r < 128; ++r) for (int c = 0; c < 128; ++c) {
                __half_raw h; h.x = (unsigned short)(o[r * 128 + c] & 0xffff); ea = fmax(ea, fabs(__half2float(__half(h)) - ref[r * 128 + c]));
                const uint32_t cell = o[r * 128 + c / 2]; h.x = (unsigned short)((c & 1) ? cell >> 16 : cell & 0xffff); eb = fmax(eb, fabs(__half2float(__half(h)) - ref[r * 128 + c]));
                const uint32_t pc = pk[r * 64 + c / 2]; h.x = (unsigned short)((c & 1) ? pc >> 16 : pc & 0xffff); ep = fmax(ep, fabs(__half2float(__half(h)) - ref[r * 128 + c]));
            }
            printf("f16: max abs err  one-per-cell(low half) %.3g | two-per-cell %.3g | pack::16b load %.3g\n", ea, eb, ep);
            printf("row 5 packed regs 0..3: %08x %08x %08x %08x\n", pk[5 * 64], pk[5 * 64 + 1], pk[5 * 64 + 2], pk[5 * 64 + 3]);
        }
    }
    return 0;
}
