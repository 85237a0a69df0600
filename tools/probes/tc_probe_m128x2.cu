// tc_probe_m128x2.cu -- where does tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA) put its accumulators in TMEM?
// One CTA pair, fp16 operands K-major without swizzle exactly as in bk_forward.cu, N = 128 (each CTA supplies 64 columns of B),
// K = 16.  MMA 1: D[row][n] = row + 1;  MMA 2 (other columns): D[row][n] = n + 1.  Every CTA dumps its 128 lanes x 256 columns.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe_m128x2 tc_probe_m128x2.cu ; run on the B200 box.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
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
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16); }
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);

// M_PER_CTA = 64 (the probe) or 128 (control: the layout bk_forward.cu relies on)
template <int M_PER_CTA>
__global__ void __cluster_dims__(2, 1, 1) probe(float *D)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    // A1, A2: [2 k-chunks][M_PER_CTA rows][8] fp16;  B1, B2: [2 k-chunks][64 cols][8]
    constexpr int A_BYTES = 2 * M_PER_CTA * 16, B_BYTES = 2 * 64 * 16;
    uint8_t *sA1 = smem, *sA2 = smem + A_BYTES, *sB1 = smem + 2 * A_BYTES, *sB2 = sB1 + B_BYTES;
    const uint32_t s_bar = smem_u32(smem + 2 * A_BYTES + 2 * B_BYTES), s_tm = s_bar + 8;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (2 * A_BYTES + 2 * B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    __syncthreads();
    for (int r = tid; r < M_PER_CTA; r += blockDim.x) {
        reinterpret_cast<__half *>(sA1 + r * 16)[0] = __float2half((float)(M_PER_CTA * rank + r + 1));     // k = 0 of chunk 0
        reinterpret_cast<__half *>(sA2 + r * 16)[0] = __float2half(1.0f);
    }
    for (int n = tid; n < 64; n += blockDim.x) {
        reinterpret_cast<__half *>(sB1 + n * 16)[0] = __float2half(1.0f);
        reinterpret_cast<__half *>(sB2 + n * 16)[0] = __float2half((float)(64 * rank + n + 1));
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s_tm) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + 2 * A_BYTES + 2 * B_BYTES + 8);
    if (rank == 0 && tid == 0) {
        constexpr uint32_t IDESC = (1u << 4) | ((128u >> 3) << 17) | (((uint32_t)(2 * M_PER_CTA) >> 4) << 24);
        const uint64_t a1 = ((uint64_t)DESC_HI << 32) | desc_lo(smem_u32(sA1), M_PER_CTA * 16);
        const uint64_t a2 = ((uint64_t)DESC_HI << 32) | desc_lo(smem_u32(sA2), M_PER_CTA * 16);
        const uint64_t b1 = ((uint64_t)DESC_HI << 32) | desc_lo(smem_u32(sB1), 1024);
        const uint64_t b2 = ((uint64_t)DESC_HI << 32) | desc_lo(smem_u32(sB2), 1024);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(a1), "l"(b1), "r"(IDESC) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem + 128u), "l"(a2), "l"(b2), "r"(IDESC) : "memory");
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(s_bar), "h"((uint16_t)3) : "memory");
    }
    long long t0 = clock64();
    while (!mbar_try_wait(s_bar, 0)) { if (clock64() - t0 > 2000000000LL) { asm volatile("trap;"); } }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        const uint32_t t_lane = tmem + ((uint32_t)(32 * warp) << 16);
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(t_lane + (uint32_t)c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 32; ++i) D[((size_t)rank * 128 + tid) * 256 + c0 + i] = __uint_as_float(v[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

template <int M>
static void run(const char *name)
{
    float *d;
    cudaMalloc(&d, 2 * 128 * 256 * 4);
    cudaMemset(d, 0xFF, 2 * 128 * 256 * 4);          // NaN pattern = never written
    const int smem = 2 * (2 * M * 16) + 2 * (2 * 64 * 16) + 64;
    cudaFuncSetAttribute(probe<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<M><<<2, 128, smem>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("== %s: %s\n", name, cudaGetErrorString(e));
    if (e != cudaSuccess) exit(1);
    float *h = (float *)malloc(2 * 128 * 256 * 4);
    cudaMemcpy(h, d, 2 * 128 * 256 * 4, cudaMemcpyDeviceToHost);
    for (int cta = 0; cta < 2; ++cta) {
        printf("CTA %d, MMA 1 (value = row + 1), column 0 and column 100 per lane:\n", cta);
        for (int lane = 0; lane < 128; ++lane) {
            const float a = h[((size_t)cta * 128 + lane) * 256 + 0], b = h[((size_t)cta * 128 + lane) * 256 + 100];
            printf("%s%3d:%g/%g", lane % 8 ? "  " : "\n  ", lane, a, b);
        }
        printf("\nCTA %d, MMA 2 (value = n + 1), lanes 0, 16, 32, 64, 96 over columns 128..255 (every 8th):\n", cta);
        const int lanes[5] = {0, 16, 32, 64, 96};
        for (int li = 0; li < 5; ++li) {
            printf("  lane %3d:", lanes[li]);
            for (int c = 128; c < 256; c += 8) printf(" %g", h[((size_t)cta * 128 + lanes[li]) * 256 + c]);
            printf("\n");
        }
    }
    free(h);
    cudaFree(d);
}

int main()
{
    run<128>("control: M = 256 (128 rows per CTA)");
    run<64>("probe: M = 128 (64 rows per CTA)");
    return 0;
}
