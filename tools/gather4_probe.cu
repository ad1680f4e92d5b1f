// Probe of cp.async.bulk.tensor.2d.tile::gather4 on sm_100a: which tensor-map box shape it wants, what lands where,
// and how many requests per cycle one SM's TMA unit accepts (tools/README.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gather4_probe gather4_probe.cu && ./gather4_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int COLS>
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, const int *__restrict__ idx, int requests,
                                             int rounds, float *__restrict__ out, long long *__restrict__ cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int *my = idx + (size_t)blockIdx.x * requests * 4;
    long long t_issue = 0, t_total = 0;
    uint32_t phase = 0;
    for (int r = 0; r < rounds; ++r) {
        if (tid == 0) {
            const long long t0 = clock64();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(requests * 4 * COLS * 4) : "memory");
            for (int q0 = 0; q0 < requests; q0 += 8) {           // row indices of 8 requests in registers before the first issue
                int4 rows[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) rows[j] = __ldg(reinterpret_cast<const int4 *>(my) + min(q0 + j, requests - 1));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int q = q0 + j;
                    if (q < requests)
                        asm volatile(
                            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                            ::"r"(smem_u32(smem + (size_t)q * 4 * COLS * 4)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(0), "r"(rows[j].x),
                              "r"(rows[j].y), "r"(rows[j].z), "r"(rows[j].w), "r"(smem_u32(&bar)) : "memory");
                }
            }
            const long long t1 = clock64();
            uint32_t ok = 0;
            int spins = 0;
            while (!ok && ++spins < (1 << 22))
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            const long long t2 = clock64();
            if (!ok) { printf("block %d: gather4 never completed\n", blockIdx.x); __trap(); }
            t_issue += t1 - t0;
            t_total += t2 - t0;
        }
        phase ^= 1;
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        const float *s = reinterpret_cast<const float *>(smem);
        for (int i = tid; i < requests * 4 * COLS; i += blockDim.x) out[i] = s[i];
        if (tid == 0) { cycles[0] = t_issue / rounds; cycles[1] = t_total / rounds; }
    }
}

int main(int argc, char **argv) {
    const int COLS = 32, N = 65536, requests = argc > 1 ? atoi(argv[1]) : 64, rounds = 20;
    const int box_rows = argc > 2 ? atoi(argv[2]) : 1;
    const int swz = argc > 3 ? atoi(argv[3]) : 0;
    const int blocks = argc > 4 ? atoi(argv[4]) : 148;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    std::vector<float> h((size_t)N * COLS);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i / COLS) + 0.001f * (float)(i % COLS);
    std::vector<int> hi((size_t)blocks * requests * 4);
    srand(1);
    for (auto &v : hi) v = rand() % N;
    float *d, *out;
    int *di;
    long long *cyc;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, (size_t)requests * 4 * COLS * 4);
    cudaMalloc(&di, hi.size() * 4);
    cudaMalloc(&cyc, 16);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(di, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {COLS, N};
    const cuuint64_t gstride[1] = {COLS * 4};
    const cuuint32_t box[2] = {COLS, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = reinterpret_cast<encode_fn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                  swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode box {%d,%d} swizzle %d: rc %d\n", COLS, box_rows, swz, (int)rc);
    if (rc != CUDA_SUCCESS) return 1;
    const size_t smem = (size_t)requests * 4 * COLS * 4 + 1024;
    cudaFuncSetAttribute(probe<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<COLS><<<blocks, 128, smem>>>(tm, di, requests, rounds, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> ho((size_t)requests * 4 * COLS);
    long long hc[2];
    cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < requests * 4 && bad < 8; ++r)
        for (int c = 0; c < COLS; ++c) {
            const float want = (float)hi[r] + 0.001f * (float)c;
            const int cc = swz ? (((c >> 2) ^ (r & 7)) << 2 | (c & 3)) : c;      // 16-byte units XOR-swizzled by the row
            if (ho[(size_t)r * COLS + cc] != want) {
                if (bad < 8) printf("row %d col %d: got %f want %f\n", r, c, ho[(size_t)r * COLS + cc], want);
                ++bad;
                break;
            }
        }
    printf("%s; %d requests of 4 x %d B per round on each of %d SMs: issue %lld cycles (%.1f per request), until landed %lld cycles\n",
           bad ? "MISMATCH" : "data ok", requests, COLS * 4, blocks, hc[0], (double)hc[0] / requests, hc[1]);
    return bad != 0;
}
