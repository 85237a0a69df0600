// tc_probe.cu -- one tcgen05.mma kind::tf32 (M = N = 128, K = 8) per variant of operand layout, checked against the host.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu ; run on the B200 box.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}

// A logical [128][8] (m, k), B logical [128][8] (n, k); a_mn / b_mn select the smem layout of each.
// K-major : offset(row, k) = (k / 4) * 2048 + row * 16 + (k % 4) * 4          SBO 128, LBO 2048
// MN-major: offset(row, k) = (row / 4) * sbo_mn + k * 16 + (row % 4) * 4      SBO sbo_mn, LBO lbo_mn
__global__ void probe(const float *A, const float *B, float *D, int a_mn, int b_mn, uint32_t sbo_mn, uint32_t lbo_mn, int swap_mn)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem, *sB = smem + 16384;
    const uint32_t s_bar = smem_u32(smem + 32768), s_tm = s_bar + 8;
    const int tid = threadIdx.x;
    for (int i = tid; i < 32768 / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 0.0f;
    __syncthreads();
    for (int i = tid; i < 128 * 8; i += blockDim.x) {
        const int row = i / 8, k = i % 8;
        const uint32_t offk = (k / 4) * 2048 + row * 16 + (k % 4) * 4;
        const uint32_t offm = (row / 4) * sbo_mn + k * 16 + (row % 4) * 4;
        *reinterpret_cast<float *>(sA + (a_mn ? offm : offk)) = A[i];
        *reinterpret_cast<float *>(sB + (b_mn ? offm : offk)) = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s_tm) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + 32768 + 8);
    if (tid == 0) {
        auto desc = [&](uint32_t addr, int mn) {
            uint32_t sbo = mn ? sbo_mn : 128u, lbo = mn ? lbo_mn : 2048u;
            if (mn && swap_mn) { uint32_t t = sbo; sbo = lbo; lbo = t; }
            const uint32_t lo = ((addr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
            const uint32_t hi = (sbo >> 4) | (1u << 14);
            return ((uint64_t)hi << 32) | lo;
        };
        const uint64_t ad = desc(smem_u32(sA), a_mn), bd = desc(smem_u32(sB), b_mn);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                               ((128u >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                     "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_bar) : "memory");
    }
    if (tid < 128) {
        long long t0 = clock64();
        while (!mbar_try_wait(s_bar, 0)) if (clock64() - t0 > 2000000000LL) asm volatile("trap;");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int warp = tid >> 5;
        for (int h = 0; h < 4; ++h) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + h * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 32; ++i) D[tid * 128 + h * 32 + i] = __uint_as_float(v[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main()
{
    const int n = 128 * 8;
    float hA[n], hB[n], *dA, *dB, *dD;
    static float hD[128 * 128], ref[128 * 128];
    srand(1);
    for (int i = 0; i < n; ++i) { hA[i] = (float)(rand() % 17 - 8) / 4.0f; hB[i] = (float)(rand() % 13 - 6) / 2.0f; }   // exact in tf32
    for (int m = 0; m < 128; ++m) for (int j = 0; j < 128; ++j) { float s = 0; for (int k = 0; k < 8; ++k) s += hA[m * 8 + k] * hB[j * 8 + k]; ref[m * 128 + j] = s; }
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    struct V { int a_mn, b_mn; uint32_t sbo, lbo; int swap; } vs[] = {
        {0, 0, 0, 0, 0}, {0, 1, 128, 128, 0}, {0, 1, 512, 128, 0}, {0, 1, 128, 512, 0}, {1, 1, 128, 128, 0}, {1, 0, 128, 128, 0},
        {0, 1, 128, 2048, 0}, {0, 1, 2048, 128, 0}};
    for (auto &v : vs) {
        cudaMemset(dD, 0xFF, sizeof hD);
        probe<<<1, 160, 40000>>>(dA, dB, dD, v.a_mn, v.b_mn, v.sbo ? v.sbo : 128, v.lbo ? v.lbo : 128, v.swap);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
        double err = 0; int nz = 0;
        for (int i = 0; i < 128 * 128; ++i) { err = fmax(err, fabs((double)hD[i] - ref[i])); nz += hD[i] != 0.0f; }
        printf("a_mn %d b_mn %d sbo_mn %u lbo_mn %u: %s max err %.4g nonzero %d  D[0][0..3] = %g %g %g %g  ref %g %g %g %g\n", v.a_mn, v.b_mn, v.sbo,
               v.lbo, cudaGetErrorString(e), err, nz, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
