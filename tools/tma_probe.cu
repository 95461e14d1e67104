// Probe: 2-D tensor-map load/store of FP64 boxes on sm_100a (which descriptor flavours the copy engine accepts).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe.bin tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int STORE>
__global__ void __launch_bounds__(32) probe(const __grid_constant__ CUtensorMap tmp, const CUtensorMap* tmg, int use_global,
                                            double* plain_out, int c0, int c1)
{
    __shared__ __align__(128) double box[8 * 32];
    __shared__ __align__(8) unsigned long long bar;
    const CUtensorMap* tm = use_global ? tmg : &tmp;
    const int lane = threadIdx.x;
    if (lane == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(&bar)), "r"(8 * 32 * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         sa(box)),
                     "l"(tm), "r"(c0), "r"(c1), "r"(sa(&bar))
                     : "memory");
    }
    __syncwarp();
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W_%=;\n}\n" ::"r"(sa(&bar)), "r"(0)
        : "memory");
    for (int k = 0; k < 8; ++k)
    {
        plain_out[k * 32 + lane] = box[k * 32 + lane];
        box[k * 32 + lane] += 1.0;
    }
    if (STORE)
    {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
        {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1),
                         "r"(sa(box))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
    }
}

int main()
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    EncodeFn encode = (EncodeFn)fn;
    const int nb = 64, rows = 62;
    std::vector<double> h((size_t)nb * 64);
    double *d, *plain;
    CUtensorMap* tmg;
    cudaMalloc(&d, h.size() * 8);
    cudaMalloc(&plain, 256 * 8);
    cudaMalloc(&tmg, sizeof(CUtensorMap));
    const CUtensorMapDataType types[2] = {CU_TENSOR_MAP_DATA_TYPE_FLOAT64, CU_TENSOR_MAP_DATA_TYPE_UINT64};
    const char* tn[2] = {"FLOAT64", "UINT64"};
    for (int ty = 0; ty < 2; ++ty)
        for (int ug = 0; ug < 2; ++ug)
            for (int st = 0; st < 2; ++st)
                for (int neg = 0; neg < 2; ++neg)
                {
                    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
                    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
                    CUtensorMap tm;
                    const cuuint64_t dims[2] = {(cuuint64_t)nb, (cuuint64_t)rows};
                    const cuuint64_t strides[1] = {(cuuint64_t)nb * 8};
                    const cuuint32_t box[2] = {32, 8};
                    const cuuint32_t es[2] = {1, 1};
                    CUresult r = encode(&tm, types[ty], 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    cudaMemcpy(tmg, &tm, sizeof tm, cudaMemcpyHostToDevice);
                    const int c0 = neg ? 48 : 32, c1 = neg ? -2 : 8;
                    if (st) probe<1><<<1, 32>>>(tm, tmg, ug, plain, c0, c1);
                    else probe<0><<<1, 32>>>(tm, tmg, ug, plain, c0, c1);
                    cudaError_t e = cudaDeviceSynchronize();
                    std::vector<double> p(256), back(h.size());
                    cudaMemcpy(p.data(), plain, 256 * 8, cudaMemcpyDeviceToHost);
                    cudaMemcpy(back.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
                    int bad_load = 0, bad_store = 0;
                    for (int k = 0; k < 8; ++k)
                        for (int l = 0; l < 32; ++l)
                        {
                            const int row = c1 + k, col = c0 + l;
                            const bool in = row >= 0 && row < rows && col < nb;
                            const double want = in ? (double)(row * nb + col) : 0.0;
                            bad_load += p[k * 32 + l] != want;
                            if (in && st) bad_store += back[row * nb + col] != want + 1.0;
                        }
                    printf("%s desc=%s store=%d coords=(%d,%d): encode=%d launch=%s bad_load=%d bad_store=%d\n", tn[ty],
                           ug ? "global" : "param", st, c0, c1, (int)r, cudaGetErrorString(e), bad_load, bad_store);
                    if (e != cudaSuccess) return 2;
                }
    return 0;
}
