// tma_copy_probe.cu -- what does the memory system deliver for the fused sweep's ACCESS PATTERN alone?
// Same grid / tiles / plane chunks / TMA boxes / 128-bit row stores as fused_sweep2_kernel, no arithmetic:
// every step waits for a plane box in a shared-memory ring, copies the tile's output region to dst, issues the box
// NB-1 steps ahead.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/tma_copy_probe tools/probe/tma_copy_probe.cu
// Run:   tools/probe/tma_copy_probe            (prints GB/s of read + written bytes per variant)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Prm {
    const float *src; float *dst;
    int pitch, rows, planes;      // floats per row, rows per plane, planes
    int LR, LGf;                  // box: rows x floats
    int OR_, OGf;                 // output rows x floats per tile (box origin = output origin - (hr, hc))
    int hr, hc;                   // halo rows / floats before the output region
    int tiles_k, tiles_j, chunk_len, NB;
    int slot_bytes;
    int do_store, do_load, order;
    int store_cs, load_hint;      // stores as st.global.cs; TMA loads with an L2 evict_first (1) / evict_last (2) policy
    int cluster_sync;             // 1: the per-step barrier is a cluster barrier (launch with cluster dims)
    int n_out_planes;
};

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 2) probe_kernel(const Prm P, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *sm = smem_dyn + ((128u - (s32(smem_dyn) & 127u)) & 127u);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm + (size_t)P.NB * P.slot_bytes);
    const int tid = threadIdx.x;
    const uint32_t mb = s32(mbar), pl = s32(sm);
    if (tid == 0) {
        for (int n = 0; n < P.NB; ++n) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb + 8u * n) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int tile, chunk;
    if (P.order == 0) { tile = blockIdx.x; chunk = blockIdx.y; }
    else { const int id = blockIdx.y * gridDim.x + blockIdx.x; chunk = id % gridDim.y; tile = id / gridDim.y; }   // chunk-fastest
    const int tk = tile % P.tiles_k, tj = tile / P.tiles_k;
    const int c0 = chunk * P.chunk_len, c1 = min(c0 + P.chunk_len, P.n_out_planes);
    if (c0 >= c1) return;
    const int R0 = tj * P.OR_, C0 = tk * P.OGf;      // box origin (storage coords; output origin = + (hr, hc))
    const int total = c1 - c0 + 4;                   // planes c0 .. c1+3 of storage (2 halo planes each side)
    const uint32_t tx = (uint32_t)P.LR * P.LGf * 4u;
    int issued = 0;
    uint64_t pol = 0;
    if (P.load_hint == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (P.load_hint == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    auto issue = [&]() {
        const uint32_t slot = issued % P.NB;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb + 8u * slot), "r"(tx) : "memory");
        if (P.load_hint)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
                             pl + slot * P.slot_bytes),
                         "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(C0), "r"(R0), "r"(c0 + issued), "r"(mb + 8u * slot), "l"(pol)
                         : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                             pl + slot * P.slot_bytes),
                         "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(C0), "r"(R0), "r"(c0 + issued), "r"(mb + 8u * slot)
                         : "memory");
        ++issued;
    };
    if (tid == 0 && P.do_load)
        for (int r = 0; r < min(P.NB - 1, total); ++r) issue();
    const int OGq = P.OGf / 4;                       // float4 groups per output row
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < total; ++s) {
        if (P.cluster_sync) {
            asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
        } else {
            __syncthreads();
        }
        if (tid == 0 && P.do_load && s + P.NB - 1 < total) issue();
        const int slot = s % P.NB;
        if (P.do_load) {
            const uint32_t par = (s / P.NB) & 1;
            uint32_t ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(mb + 8u * slot), "r"(par) : "memory");
            } while (!ok);
        }
        if (s < 2 || s >= total - 2) continue;       // halo planes are only read
        const float4 *splane = reinterpret_cast<const float4 *>(sm + (size_t)slot * P.slot_bytes);
        float *dplane = P.dst + (size_t)(c0 + s) * P.pitch * P.rows;
        for (int it = tid; it < P.OR_ * OGq; it += 256) {
            const int r = it / OGq, q = it - r * OGq;
            const int gr = R0 + P.hr + r, gc = C0 + P.hc + 4 * q;
            float4 v = P.do_load ? splane[((P.hr + r) * P.LGf + P.hc) / 4 + q] : make_float4(1.f, 2.f, 3.f, 4.f);
            if (gr < P.rows && gc + 3 < P.pitch) {
                if (P.do_store) {
                    if (P.store_cs) __stcs(reinterpret_cast<float4 *>(dplane + (size_t)gr * P.pitch + gc), v);
                    else *reinterpret_cast<float4 *>(dplane + (size_t)gr * P.pitch + gc) = v;
                }
                else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
            }
        }
    }
    if (!P.do_store && acc.x + acc.y + acc.z + acc.w == 123.456f) P.dst[0] = acc.x;
}

__global__ void copy_kernel(const float4 *__restrict__ a, float4 *__restrict__ b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

typedef CUresult (*EncFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 512;
    const int pitch = N + 8, rows = N + 4, planes = N + 4;
    const size_t n = (size_t)pitch * rows * planes;
    float *src, *dst;
    CK(cudaMalloc(&src, n * 4)); CK(cudaMalloc(&dst, n * 4));
    CK(cudaMemset(src, 0, n * 4)); CK(cudaMemset(dst, 0, n * 4));
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncFn enc = (EncFn)fp;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    {   // plain copy for reference
        for (int w = 0; w < 3; ++w) copy_kernel<<<148 * 8, 512>>>((const float4 *)src, (float4 *)dst, n / 4);
        CK(cudaEventRecord(e0));
        for (int w = 0; w < 10; ++w) copy_kernel<<<148 * 8, 512>>>((const float4 *)src, (float4 *)dst, n / 4);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("plain copy                                              : %7.1f us  %6.0f GB/s (read + write)\n", ms * 100, 2.0 * n * 4 * 10 / ms / 1e6);
    }
    struct Var { const char *name; int LR, LGq, OR_, OGq, hr, hcq, NB, promo, do_load, do_store, order, chunks, store_cs, load_hint, cx, cy, csync; };
    const Var vars[] = {
        {"fused geometry (box 30x35g, out 26x32g, NB 6), 20 chunks", 30, 35, 26, 32, 2, 1, 6, 3, 1, 1, 0, 20, 0, 0, 1, 1, 0},
        {"  cluster 4x1 (z neighbours), no cluster barrier", 30, 35, 26, 32, 2, 1, 6, 3, 1, 1, 0, 20, 0, 0, 4, 1, 0},
        {"  cluster 4x1, cluster barrier every step", 30, 35, 26, 32, 2, 1, 6, 3, 1, 1, 0, 20, 0, 0, 4, 1, 1},
        {"  cluster 2x1, cluster barrier every step", 30, 35, 26, 32, 2, 1, 6, 3, 1, 1, 0, 20, 0, 0, 2, 1, 1},
        {"  cluster 8x1 (4 z x 2 y), cluster barrier every step", 30, 35, 26, 32, 2, 1, 6, 3, 1, 1, 0, 20, 0, 0, 8, 1, 1},
        {"  loads only", 30, 35, 26, 32, 2, 1, 6, 3, 1, 0, 0, 20, 0, 0, 1, 1, 0},
        {"  loads only, cluster 4x1 + barrier", 30, 35, 26, 32, 2, 1, 6, 3, 1, 0, 0, 20, 0, 0, 4, 1, 1},
        {"  loads only, cluster 8x1 + barrier", 30, 35, 26, 32, 2, 1, 6, 3, 1, 0, 0, 20, 0, 0, 8, 1, 1},
        {"no halo: box 26x32g = out, 20 chunks", 26, 32, 26, 32, 0, 0, 6, 3, 1, 1, 0, 20, 0, 0, 1, 1, 0},
        {"no halo, loads only", 26, 32, 26, 32, 0, 0, 6, 3, 1, 0, 0, 20, 0, 0, 1, 1, 0},
    };
    for (const Var &v : vars) {
        Prm P; memset(&P, 0, sizeof(P));
        P.src = src; P.dst = dst; P.pitch = pitch; P.rows = rows; P.planes = planes;
        P.LR = v.LR; P.LGf = v.LGq * 4; P.OR_ = v.OR_; P.OGf = v.OGq * 4; P.hr = v.hr; P.hc = v.hcq * 4; P.NB = v.NB;
        P.tiles_j = (N + v.OR_ - 1) / v.OR_; P.tiles_k = (N / 4 + v.OGq - 1) / v.OGq;
        P.n_out_planes = N;
        P.chunk_len = (N + v.chunks - 1) / v.chunks; P.chunk_len += P.chunk_len & 1;
        const int chunks = (N + P.chunk_len - 1) / P.chunk_len;
        P.slot_bytes = ((v.LR * v.LGq * 16 + 127) / 128) * 128;
        P.do_load = v.do_load; P.do_store = v.do_store; P.order = v.order; P.store_cs = v.store_cs; P.load_hint = v.load_hint; P.cluster_sync = v.csync;
        const size_t smem = (size_t)P.NB * P.slot_bytes + 128 + 128;
        if (smem > 113 * 1024) { printf("%-56s: skipped (%zu B of shared memory)\n", v.name, smem); continue; }
        CUtensorMap tm;
        const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)planes};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * rows * 4};
        const cuuint32_t box[3] = {(cuuint32_t)P.LGf, (cuuint32_t)P.LR, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)v.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-56s: encode failed %d\n", v.name, (int)r); continue; }
        CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(P.tiles_j * P.tiles_k, chunks);
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = grid; cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = v.cx; at[0].val.clusterDim.y = v.cy; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = (v.cx * v.cy > 1) ? 1 : 0;
        if (grid.x % v.cx || grid.y % v.cy) { printf("%-56s: grid %dx%d not divisible by the cluster\n", v.name, grid.x, grid.y); continue; }
        for (int w = 0; w < 3; ++w) CK(cudaLaunchKernelEx(&cfg, probe_kernel, P, tm));
        CK(cudaGetLastError());
        CK(cudaEventRecord(e0));
        const int reps = 20;
        for (int w = 0; w < reps; ++w) CK(cudaLaunchKernelEx(&cfg, probe_kernel, P, tm));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double useful = (double)N * N * N * 4 * ((v.do_load ? 1 : 0) + (v.do_store ? 1 : 0));
        const double req = (double)grid.x * chunks * (P.chunk_len + 4) * v.LR * v.LGq * 16 * (v.do_load ? 1 : 0) +
                           (double)N * N * N * 4 * (v.do_store ? 1 : 0);
        printf("%-56s: %7.1f us  useful %6.0f GB/s  requested %6.0f GB/s  grid %dx%d smem %zu\n", v.name, ms / reps * 1e3,
               useful * reps / ms / 1e6, req * reps / ms / 1e6, grid.x, chunks, smem);
    }
    return 0;
}
