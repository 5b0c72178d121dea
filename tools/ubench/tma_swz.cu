// Where does a TMA tile copy with a 32-byte inner box put its rows under SWIZZLE_128B?  (dev probe)
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_swz tma_swz.cu && ./tma_swz
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, unsigned long long* out, int x0, int words)
{
    extern __shared__ __align__(1024) unsigned char raw[];
    __shared__ __align__(8) unsigned long long bar;
    unsigned long long* sm = reinterpret_cast<unsigned long long*>(raw);
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = 0xdeadbeefdeadbeefull;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar), dst = (unsigned)__cvta_generic_to_shared(raw);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(64 * 32) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(&map), "r"(x0), "r"(0), "r"(b) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < words; i += blockDim.x) out[i] = sm[i];
}
int main()
{
    const int rows = 64, cols = 512;
    std::vector<unsigned long long> h((size_t)rows * cols);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) h[(size_t)r * cols + c] = ((unsigned long long)r << 16) | c;
    unsigned long long *d, *o;
    cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    const int words = 2048;                       // 16 KB of shared memory observed
    cudaMalloc(&o, words * 8);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    for (int mode = 0; mode < 2; ++mode) for (int x0 = 0; x0 <= 4; x0 += 4) {
        const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)cols * 8};
        const cuuint32_t box[2] = {4, 64}, es[2] = {1, 1};
        CUtensorMap m;
        CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          mode ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("mode %s x0 %d encode rc %d\n", mode ? "SW128" : "NONE", x0, (int)rc);
        if (rc) continue;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 8);
        probe<<<1, 128, words * 8>>>(m, o, x0, words);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  run: %s\n", cudaGetErrorString(e));
        if (e) return 1;
        std::vector<unsigned long long> r(words);
        cudaMemcpy(r.data(), o, words * 8, cudaMemcpyDeviceToHost);
        int last = -1;
        for (int i = 0; i < words; ++i) if (r[i] != 0xdeadbeefdeadbeefull) last = i;
        printf("  last written word %d (%d bytes)\n", last, (last + 1) * 8);
        for (int i = 0; i < 64 && i * 4 <= last; ++i) {     // 32-byte slots: which (row, col) landed there
            const unsigned long long v = r[i * 4];
            if (v == 0xdeadbeefdeadbeefull) printf("  slot %2d: -\n", i);
            else printf("  slot %2d: row %llu col %llu\n", i, v >> 16, v & 0xffff);
            if (i == 17) break;
        }
    }
    return 0;
}
