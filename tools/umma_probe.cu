// umma_probe.cu -- stand-alone probe of tcgen05.mma shared-memory descriptor conventions (no-swizzle layouts).
// Fills shared memory with small integers (exact in tf32 / bf16), issues ONE MMA (M=128, N=32) with the requested
// major-ness / LBO / SBO, reads D back from TMEM and compares it with the host model
//     off(mn, k) = (mn % EPC) * ELEM + (mn / EPC) * MNSTRIDE + (k % 8) * 16 + (k / 8) * KSTRIDE      (MN-major)
//     off(mn, k) = (k  % EPC) * ELEM + (k  / EPC) * KSTRIDE  + (mn % 8) * 16 + (mn / 8) * MNSTRIDE   (K-major)
// under both assignments of {LBO, SBO} to {MNSTRIDE, KSTRIDE}.    build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#include "../tec_mollm_b200/csrc/tc.cuh"

using namespace tg;

constexpr int kSmemBytes = 96 * 1024;

struct ProbeArgs {
    int bf16, a_mn, b_mn;
    uint32_t a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo;
    int N;
};

__global__ void __launch_bounds__(128) probe_kernel(const uint32_t *init, float *d_out, ProbeArgs p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kSmemBytes / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = init[i];
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_ptr, 64);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_ptr;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(p.bf16 ? kFmtBF16 : kFmtTF32, 128, p.N, p.a_mn, p.b_mn);
        const uint64_t da = umma_desc(smem_u32(smem) + p.a_off, p.a_lbo, p.a_sbo);
        const uint64_t db = umma_desc(smem_u32(smem) + p.b_off, p.b_lbo, p.b_sbo);
        if (p.bf16) umma_bf16(tbase, da, db, idesc, 0);
        else umma_tf32(tbase, da, db, idesc, 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int cb = 0; cb < p.N / 16; ++cb) {
        float v[16];
        tmem_ld16(taddr + cb * 16, v);
        for (int i = 0; i < 16; ++i) d_out[(warp * 32 + lane) * p.N + cb * 16 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 64);
}

static std::vector<uint32_t> g_init;
static float elem_at(uint32_t byte_off, int bf16) {
    if (byte_off + 4 > (uint32_t)kSmemBytes) return NAN;
    if (bf16) {
        uint16_t h;
        memcpy(&h, reinterpret_cast<const unsigned char *>(g_init.data()) + byte_off, 2);
        uint32_t w = (uint32_t)h << 16;
        float f;
        memcpy(&f, &w, 4);
        return f;
    }
    float f;
    memcpy(&f, reinterpret_cast<const unsigned char *>(g_init.data()) + byte_off, 4);
    return f;
}
static uint32_t off_model(int mn_major, int mn, int k, uint32_t mnstride, uint32_t kstride, int bf16) {
    const int elem = bf16 ? 2 : 4, epc = 16 / elem;
    if (mn_major) return (mn % epc) * elem + (mn / epc) * mnstride + (k % 8) * 16 + (k / 8) * kstride;
    return (k % epc) * elem + (k / epc) * kstride + (mn % 8) * 16 + (mn / 8) * mnstride;
}

int main() {
    g_init.resize(kSmemBytes / 4);
    // small integers, exact in tf32 and (as two packed halves) in bf16
    for (size_t i = 0; i < g_init.size(); ++i) {
        float f = (float)((int)((i * 2654435761u >> 7) % 15) - 7);
        memcpy(&g_init[i], &f, 4);
    }
    uint32_t *d_init;
    float *d_out;
    cudaMalloc(&d_init, kSmemBytes);
    cudaMalloc(&d_out, 128 * 256 * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    const int N = 32;
    std::vector<float> out(128 * N);
    struct Case { const char *name; int bf16, a_mn, b_mn; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; };
    const uint32_t PA = 1536, PB = 1024;  // distinct, so the two hypotheses give different answers
    Case cases[] = {
        {"tf32 K/K    lbo=128 sbo=P", 0, 0, 0, 128, PA, 128, PB},
        {"tf32 MN/K   A: lbo=P sbo=128", 0, 1, 0, PA, 128, 128, PB},
        {"tf32 MN/K   A: lbo=128 sbo=P", 0, 1, 0, 128, PA, 128, PB},
        {"tf32 K/MN   B: lbo=P sbo=128", 0, 0, 1, 128, PA, PB, 128},
        {"tf32 K/MN   B: lbo=128 sbo=P", 0, 0, 1, 128, PA, 128, PB},
        {"tf32 MN/MN  lbo=P sbo=128", 0, 1, 1, PA, 128, PB, 128},
        {"bf16 K/K    lbo=128 sbo=P", 1, 0, 0, 128, PA, 128, PB},
        {"bf16 MN/MN  lbo=P sbo=128", 1, 1, 1, PA, 128, PB, 128},
        {"bf16 MN/MN  lbo=128 sbo=P", 1, 1, 1, 128, PA, 128, PB},
    };
    for (const Case &c : cases) {
        if (c.bf16) {  // refill with bf16-exact halves
            for (size_t i = 0; i < g_init.size(); ++i) {
                float lo = (float)((int)((i * 2654435761u >> 7) % 15) - 7), hi = (float)((int)((i * 40503u >> 3) % 13) - 6);
                uint32_t wl, wh;
                memcpy(&wl, &lo, 4);
                memcpy(&wh, &hi, 4);
                g_init[i] = (wl >> 16) | (wh & 0xFFFF0000u);
            }
        } else {
            for (size_t i = 0; i < g_init.size(); ++i) {
                float f = (float)((int)((i * 2654435761u >> 7) % 15) - 7);
                memcpy(&g_init[i], &f, 4);
            }
        }
        cudaMemcpy(d_init, g_init.data(), kSmemBytes, cudaMemcpyHostToDevice);
        ProbeArgs p{c.bf16, c.a_mn, c.b_mn, 0u, 48u * 1024u, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, N};
        cudaMemset(d_out, 0xFF, 128 * N * 4);
        probe_kernel<<<1, 128, kSmemBytes>>>(d_init, d_out, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%-34s CUDA error: %s\n", c.name, cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(out.data(), d_out, 128 * N * 4, cudaMemcpyDeviceToHost);
        const int K = c.bf16 ? 16 : 8;
        int zeros = 0;
        for (float v : out) zeros += (v == 0.f);
        printf("%-34s zeros %4d/%d |", c.name, zeros, 128 * N);
        // hypotheses: which of (lbo, sbo) is the MN stride, independently for A and B
        for (int ha = 0; ha < 2; ++ha)
            for (int hb = 0; hb < 2; ++hb) {
                const uint32_t a_mns = ha ? c.a_lbo : c.a_sbo, a_ks = ha ? c.a_sbo : c.a_lbo;
                const uint32_t b_mns = hb ? c.b_lbo : c.b_sbo, b_ks = hb ? c.b_sbo : c.b_lbo;
                double maxd = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < N; ++n) {
                        double acc = 0;
                        for (int k = 0; k < K; ++k)
                            acc += (double)elem_at(p.a_off + off_model(c.a_mn, m, k, a_mns, a_ks, c.bf16), c.bf16) *
                                   (double)elem_at(p.b_off + off_model(c.b_mn, n, k, b_mns, b_ks, c.bf16), c.bf16);
                        maxd = fmax(maxd, fabs(acc - (double)out[m * N + n]));
                    }
                printf(" A.mn=%s B.mn=%s: %8.1f |", ha ? "LBO" : "SBO", hb ? "LBO" : "SBO", maxd);
            }
        printf("\n");
    }
    return 0;
}
