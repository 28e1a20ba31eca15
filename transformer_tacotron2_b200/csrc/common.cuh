// Shared device helpers for the sm_100a Transformer-TTS kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>
#include <mutex>

typedef __nv_bfloat16 bf16;

#define TTS_HD __host__ __device__ __forceinline__
#define TTS_D __device__ __forceinline__

namespace tts {

// Kernel launches issued by this library since load (bench.py reports it as gpu_launches).
inline unsigned long long& launch_counter() { static unsigned long long n = 0; return n; }

// One-time per-DEVICE launcher state (cudaFuncSetAttribute and the SM count are per device; a process may hold handles on
// several GPUs).  `setup` runs once for the calling thread's current device; *num_sms receives that device's SM count.
struct PerDevice {
    std::mutex mu;
    bool done[64] = {};
    int sms[64] = {};
};
template <class Fn>
inline cudaError_t per_device_once(PerDevice& s, int* num_sms, Fn setup) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(s.mu);
    if (!s.done[dev]) {
        if ((e = setup()) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&s.sms[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        s.done[dev] = true;
    }
    if (num_sms) *num_sms = s.sms[dev];
    return cudaSuccess;
}
// Makes `dev` current for the lifetime of the guard and restores the caller's device afterwards (the C ABI must not
// change the current device under PyTorch's feet).
struct DeviceGuard {
    int prev = -1, target;
    cudaError_t err;
    explicit DeviceGuard(int dev) : target(dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0 && prev != target) cudaSetDevice(prev); }
};

constexpr int kDModel = 512;     // the base model is the only model on this path
constexpr int kHeads = 8;
constexpr int kDHead = 64;
constexpr float kLog2e = 1.4426950408889634f;

// ---------------------------------------------------------------- small utilities
TTS_D uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
TTS_D float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
// 2^x on the SFU (ex2.approx): -inf -> 0, exact enough for softmax weights that are rounded to bf16 anyway
TTS_D float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
TTS_D float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
TTS_D float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// L2 eviction policies (createpolicy): KV-cache rows are streamed once per step -> evict_first, so
// they do not push the weights (re-read every step, 44.7 MB) out of the 126 MB L2 -> evict_last.
TTS_D uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
TTS_D uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// ---------------------------------------------------------------- warp-level bf16 MMA (m16n8k16)
// Decode K/V cache block (decode_cluster.cuh): 64 rows of one (layer, utterance, head) = 8192 bf16 = four 16-row sub-chunks of
// 2048 elements [K 16 rows row-major [16][64] | V 16 rows in mma.m16n8k16 A-FRAGMENT order of V^T [64 d][16 rows]]: for the tile
// dt = d / 16 lane (g = d % 8, t4 = (r % 16) / 4) holds {V[4 t4 + 0..1][g], V[4 t4 + 0..1][g + 8], V[4 t4 + 2..3][g], V[4 t4 + 2..3][g + 8]}
// as one 16-byte chunk, so the attention loop fetches a whole A fragment with one LDS.128.  Any run of rows [0, 16 n) of a
// (layer, utterance, head) is one contiguous range of 16 n * 256 bytes: a ring stage is always ONE bulk copy.
__host__ __device__ __forceinline__ int kv_k_elem(int r, int d) { return (r >> 4) * 2048 + (r & 15) * 64 + d; }
__host__ __device__ __forceinline__ int kv_v_elem(int r, int d) {
    const int rr = r & 15;
    return (r >> 4) * 2048 + 1024 + (((d >> 4) * 32 + (d & 7) * 4 + (rr >> 2)) << 3) + ((((rr >> 1) & 1) * 2 + ((d >> 3) & 1)) << 1) + (rr & 1);
}

TTS_D void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
TTS_D void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
TTS_D void ldmatrix_x2(uint32_t& r0, uint32_t& r1, const void* smem_ptr) {
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
TTS_D void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
TTS_D void cp_async_16(void* smem_ptr, const void* gptr, bool valid) {
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    int sz = valid ? 16 : 0;                       // src-size 0 => 16 bytes of zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(addr), "l"(gptr), "r"(sz));
}
TTS_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
TTS_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

TTS_D float to_f32(float x) { return x; }
TTS_D float to_f32(bf16 x) { return __bfloat162float(x); }

}  // namespace tts
