// wbench.cu -- write-bandwidth micro-benchmarks used to choose the render kernel's store pattern.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/wbench tools/wbench.cu
//   tools/wbench [GiB]            (GPU box only)
//
// Each variant writes the same big buffer (default 96 GiB) and reports GB/s.  The question it
// answers: which (engine, chunk size, CTA->address order, cache policy) reaches the pure-write
// rate that a trivial fill kernel gets on this GPU.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

constexpr uint32_t ENV_BYTES = 112896;   // one v0 observation

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int POLICY>
__device__ __forceinline__ void st16(void *p, const uint4 &v) {
  if (POLICY == 0)
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else if (POLICY == 1)
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else if (POLICY == 2)
    asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else
    asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// K1: torch-like fill: small CTAs, each writes 128 threads x 4 x 16 B = 8 KB, block order = address order
template <int POLICY>
__global__ void k_fill_small(uint4 *buf, size_t n16) {
  size_t base = (size_t)blockIdx.x * 512 + threadIdx.x;
  uint4 v = make_uint4(0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    size_t i = base + k * 128;
    if (i < n16) st16<POLICY>(buf + i, v);
  }
}

// K2: persistent fill: grid CTAs walk the buffer together, chunk = THREADS*16*4 bytes
template <int POLICY, int THREADS>
__global__ void __launch_bounds__(THREADS) k_fill_persist(uint4 *buf, size_t n16) {
  uint4 v = make_uint4(0x3f800000u, 0, 0x3f800000u, 0);
  const size_t chunk = (size_t)THREADS * 4;
  for (size_t c = blockIdx.x; c * chunk < n16; c += gridDim.x) {
    size_t base = c * chunk + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      size_t i = base + (size_t)k * THREADS;
      if (i < n16) st16<POLICY>(buf + i, v);
    }
  }
}

// K3: persistent LDS->STG copy of env-sized images from a shared-memory template.
// PARTS = pieces an env is split into; piece c of the flat piece sequence goes to CTA c % grid.
template <int POLICY, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_smem_copy(unsigned char *buf, size_t nenv, int parts) {
  extern __shared__ __align__(128) unsigned char sm[];
  for (uint32_t i = threadIdx.x; i < ENV_BYTES / 4; i += THREADS) ((uint32_t *)sm)[i] = (i * 2654435761u) >> 31 ? 0x3f800000u : 0u;
  __syncthreads();
  const uint32_t piece = ENV_BYTES / parts;            // multiple of 16 for parts in {1,2,3,4,6,7,8,12,14,21,24}
  const uint32_t p16 = piece / 16;
  const size_t npieces = nenv * parts;
  for (size_t c = blockIdx.x; c < npieces; c += gridDim.x) {
    const uint32_t part = (uint32_t)(c % parts);
    unsigned char *dst = buf + c * piece;
    const uint32_t src = smem_addr(sm) + part * piece;
    uint32_t f = threadIdx.x;
    for (; f + 3 * THREADS < p16; f += 4 * THREADS) {
      uint4 v0, v1, v2, v3;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0.x), "=r"(v0.y), "=r"(v0.z), "=r"(v0.w) : "r"(src + (f << 4)));
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v1.x), "=r"(v1.y), "=r"(v1.z), "=r"(v1.w) : "r"(src + ((f + THREADS) << 4)));
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v2.x), "=r"(v2.y), "=r"(v2.z), "=r"(v2.w) : "r"(src + ((f + 2 * THREADS) << 4)));
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v3.x), "=r"(v3.y), "=r"(v3.z), "=r"(v3.w) : "r"(src + ((f + 3 * THREADS) << 4)));
      st16<POLICY>(dst + ((size_t)f << 4), v0);
      st16<POLICY>(dst + ((size_t)(f + THREADS) << 4), v1);
      st16<POLICY>(dst + ((size_t)(f + 2 * THREADS) << 4), v2);
      st16<POLICY>(dst + ((size_t)(f + 3 * THREADS) << 4), v3);
    }
    for (; f < p16; f += THREADS) {
      uint4 v;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src + (f << 4)));
      st16<POLICY>(dst + ((size_t)f << 4), v);
    }
  }
}

// K4: persistent TMA bulk shared->global copies.  order 0: piece c -> CTA c % grid (fill-like);
// order 1: CTA b owns a contiguous range of pieces.  ISSUERS threads per CTA issue round-robin.
template <int ISSUERS>
__global__ void __launch_bounds__(ISSUERS * 32, 1) k_tma_copy(unsigned char *buf, size_t nenv, int parts, int order) {
  extern __shared__ __align__(128) unsigned char sm[];
  for (uint32_t i = threadIdx.x; i < ENV_BYTES / 4; i += blockDim.x) ((uint32_t *)sm)[i] = (i * 2654435761u) >> 31 ? 0x3f800000u : 0u;
  fence_async();
  __syncthreads();
  if ((threadIdx.x & 31) != 0) return;
  const int w = threadIdx.x >> 5;
  const uint32_t piece = ENV_BYTES / parts;
  const size_t npieces = nenv * parts;
  const uint32_t sbase = smem_addr(sm);
  if (order == 0) {
    for (size_t c = (size_t)w * gridDim.x + blockIdx.x; c < npieces; c += (size_t)gridDim.x * ISSUERS)
      bulk_s2g(buf + c * piece, sbase + (uint32_t)(c % parts) * piece, piece);
  } else {
    const size_t lo = npieces * blockIdx.x / gridDim.x, hi = npieces * (blockIdx.x + 1) / gridDim.x;
    for (size_t c = lo + w; c < hi; c += ISSUERS)
      bulk_s2g(buf + c * piece, sbase + (uint32_t)(c % parts) * piece, piece);
  }
  bulk_commit();
  bulk_wait_all();
}


__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// K2d: persistent fill, chunks handed out by an atomic counter (dynamic load balance)
template <int POLICY, int THREADS>
__global__ void __launch_bounds__(THREADS) k_fill_dynamic(uint4 *buf, size_t n16, unsigned long long *counter, uint32_t chunk16) {
  __shared__ unsigned long long s_c;
  uint4 v = make_uint4(0x3f800000u, 0, 0x3f800000u, 0);
  const size_t nchunks = (n16 + chunk16 - 1) / chunk16;
  for (;;) {
    if (threadIdx.x == 0) s_c = atomicAdd(counter, 1ull);
    __syncthreads();
    const size_t c = s_c;
    __syncthreads();
    if (c >= nchunks) break;
    const size_t lo = c * chunk16, hi = (lo + chunk16 < n16) ? lo + chunk16 : n16;
    for (size_t i = lo + threadIdx.x; i < hi; i += THREADS) st16<POLICY>(buf + i, v);
  }
}

// K4d: TMA copies, work items (envs_per_grab envs, each split in `parts`) handed out by an atomic
// counter; every warp of the CTA is an independent issuer.  finish[] gets each CTA's end time.
template <int ISSUERS>
__global__ void __launch_bounds__(ISSUERS * 32, 1) k_tma_dynamic(unsigned char *buf, size_t nenv, int parts, int envs_per_grab,
                                                                unsigned long long *counter, unsigned long long *finish) {
  extern __shared__ __align__(128) unsigned char sm[];
  for (uint32_t i = threadIdx.x; i < ENV_BYTES / 4; i += blockDim.x) ((uint32_t *)sm)[i] = (i * 2654435761u) >> 31 ? 0x3f800000u : 0u;
  fence_async();
  __syncthreads();
  if ((threadIdx.x & 31) != 0) return;
  const uint32_t piece = ENV_BYTES / parts;
  const uint32_t sbase = smem_addr(sm);
  const size_t ngrabs = (nenv + envs_per_grab - 1) / envs_per_grab;
  for (;;) {
    const size_t g = counter ? (size_t)atomicAdd(counter, 1ull) : 0;
    if (g >= ngrabs) break;
    const size_t e0 = g * envs_per_grab, e1 = (e0 + envs_per_grab < nenv) ? e0 + envs_per_grab : nenv;
    for (size_t e = e0; e < e1; ++e) {
      unsigned char *dst = buf + e * ENV_BYTES;
      for (int q = 0; q < parts; ++q) bulk_s2g(dst + (size_t)q * piece, sbase + q * piece, piece);
    }
  }
  bulk_commit();
  bulk_wait_all();
  if (finish && threadIdx.x == 0) finish[blockIdx.x] = gtime();
}

// K4s: static round-robin (env e -> CTA e % grid) with finish times, to look for straggler SMs
__global__ void __launch_bounds__(32, 1) k_tma_static_timed(unsigned char *buf, size_t nenv, unsigned long long *finish) {
  extern __shared__ __align__(128) unsigned char sm[];
  for (uint32_t i = threadIdx.x; i < ENV_BYTES / 4; i += blockDim.x) ((uint32_t *)sm)[i] = (i * 2654435761u) >> 31 ? 0x3f800000u : 0u;
  fence_async();
  __syncthreads();
  if (threadIdx.x != 0) return;
  const uint32_t sbase = smem_addr(sm);
  for (size_t e = blockIdx.x; e < nenv; e += gridDim.x) bulk_s2g(buf + e * ENV_BYTES, sbase, ENV_BYTES);
  bulk_commit();
  bulk_wait_all();
  finish[blockIdx.x] = gtime();
}

struct Timer {
  cudaEvent_t a, b;
  Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
  template <class F>
  float run(F f, int reps = 3) {
    f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
  }
};

int main(int argc, char **argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 96.0;
  const size_t nenv = (size_t)(gib * (1ull << 30) / ENV_BYTES);
  const size_t bytes = nenv * ENV_BYTES, n16 = bytes / 16;
  unsigned char *buf;
  CK(cudaMalloc(&buf, bytes));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("# %s, %d SMs, buffer %.1f GB (%zu envs of %u B)\n", prop.name, sms, bytes / 1e9, nenv, ENV_BYTES);
  Timer t;
  auto rep = [&](const char *name, float ms) { printf("%-64s %8.3f ms %8.1f GB/s\n", name, ms, bytes / ms / 1e6); fflush(stdout); };

  rep("cudaMemsetAsync", t.run([&] { CK(cudaMemsetAsync(buf, 0, bytes)); }));
  rep("K1 fill small CTAs (8 KB/CTA, st default)", t.run([&] { k_fill_small<0><<<(unsigned)((n16 + 511) / 512), 128>>>((uint4 *)buf, n16); }));
  rep("K1 fill small CTAs (st.cs)", t.run([&] { k_fill_small<1><<<(unsigned)((n16 + 511) / 512), 128>>>((uint4 *)buf, n16); }));
  rep("K2 fill persistent 148x1024 (st default)", t.run([&] { k_fill_persist<0, 1024><<<sms, 1024>>>((uint4 *)buf, n16); }));
  rep("K2 fill persistent 148x1024 (st.cs)", t.run([&] { k_fill_persist<1, 1024><<<sms, 1024>>>((uint4 *)buf, n16); }));
  rep("K2 fill persistent 148x1024 (st.cg)", t.run([&] { k_fill_persist<2, 1024><<<sms, 1024>>>((uint4 *)buf, n16); }));
  rep("K2 fill persistent 148x1024 (st.wt)", t.run([&] { k_fill_persist<3, 1024><<<sms, 1024>>>((uint4 *)buf, n16); }));
  rep("K2 fill persistent 296x512 (st default)", t.run([&] { k_fill_persist<0, 512><<<2 * sms, 512>>>((uint4 *)buf, n16); }));
  rep("K2 fill persistent 592x256 (st default)", t.run([&] { k_fill_persist<0, 256><<<4 * sms, 256>>>((uint4 *)buf, n16); }));
  rep("K2 fill persistent 1184x256 (st default)", t.run([&] { k_fill_persist<0, 256><<<8 * sms, 256>>>((uint4 *)buf, n16); }));

  CK(cudaFuncSetAttribute(k_smem_copy<0, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENV_BYTES));
  CK(cudaFuncSetAttribute(k_smem_copy<1, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENV_BYTES));
  CK(cudaFuncSetAttribute(k_smem_copy<0, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENV_BYTES));
  for (int parts : {1, 4, 24}) {
    char nm[128];
    snprintf(nm, sizeof nm, "K3 smem->STG 148x1024, env split in %d (st default)", parts);
    rep(nm, t.run([&] { k_smem_copy<0, 1024><<<sms, 1024, ENV_BYTES>>>(buf, nenv, parts); }));
    snprintf(nm, sizeof nm, "K3 smem->STG 148x1024, env split in %d (st.cs)", parts);
    rep(nm, t.run([&] { k_smem_copy<1, 1024><<<sms, 1024, ENV_BYTES>>>(buf, nenv, parts); }));
  }
  rep("K3 smem->STG 296x512 (2 CTA/SM), split 24 (st default)",
      t.run([&] { k_smem_copy<0, 512><<<2 * sms, 512, ENV_BYTES>>>(buf, nenv, 24); }));

  CK(cudaFuncSetAttribute(k_tma_copy<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENV_BYTES));
  CK(cudaFuncSetAttribute(k_tma_copy<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENV_BYTES));
  for (int order : {0, 1})
    for (int parts : {1, 4, 24, 84}) {
      char nm[128];
      snprintf(nm, sizeof nm, "K4 TMA S2G 1 issuer/SM, split %d, order %s", parts, order ? "chunk-per-CTA" : "round-robin");
      rep(nm, t.run([&] { k_tma_copy<1><<<sms, 32, ENV_BYTES>>>(buf, nenv, parts, order); }));
    }
  for (int parts : {1, 24}) {
    char nm[128];
    snprintf(nm, sizeof nm, "K4 TMA S2G 4 issuers/SM, split %d, round-robin", parts);
    rep(nm, t.run([&] { k_tma_copy<4><<<sms, 128, ENV_BYTES>>>(buf, nenv, parts, 0); }));
  }
  rep("K4 TMA S2G 2 CTAs/SM x 1 issuer, split 1, round-robin",
      t.run([&] { k_tma_copy<1><<<2 * sms, 32, ENV_BYTES>>>(buf, nenv, 1, 0); }));

  // ---- dynamic work distribution
  unsigned long long *counter, *finish;
  CK(cudaMalloc(&counter, 8));
  CK(cudaMalloc(&finish, 8 * 4096));
  auto zero = [&] { CK(cudaMemsetAsync(counter, 0, 8)); };
  for (uint32_t chunk_kb : {16u, 64u, 256u}) {
    char nm[128];
    snprintf(nm, sizeof nm, "K2d fill persistent 148x1024 DYNAMIC chunks of %u KB", chunk_kb);
    rep(nm, t.run([&] { zero(); k_fill_dynamic<0, 1024><<<sms, 1024>>>((uint4 *)buf, n16, counter, chunk_kb * 64); }));
  }
  rep("K2d fill persistent 296x512 DYNAMIC chunks of 64 KB",
      t.run([&] { zero(); k_fill_dynamic<0, 512><<<2 * sms, 512>>>((uint4 *)buf, n16, counter, 64 * 64); }));
  CK(cudaFuncSetAttribute(k_tma_dynamic<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENV_BYTES));
  CK(cudaFuncSetAttribute(k_tma_dynamic<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENV_BYTES));
  CK(cudaFuncSetAttribute(k_tma_dynamic<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENV_BYTES));
  CK(cudaFuncSetAttribute(k_tma_static_timed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENV_BYTES));
  for (int parts : {1, 3, 24})
    for (int epg : {1, 8, 32}) {
      char nm[128];
      snprintf(nm, sizeof nm, "K4d TMA DYNAMIC 1 issuer/SM, %d env/grab, split %d", epg, parts);
      rep(nm, t.run([&] { zero(); k_tma_dynamic<1><<<sms, 32, ENV_BYTES>>>(buf, nenv, parts, epg, counter, nullptr); }));
    }
  for (int parts : {1, 24}) {
    char nm[128];
    snprintf(nm, sizeof nm, "K4d TMA DYNAMIC 2 issuers/SM, 8 env/grab, split %d", parts);
    rep(nm, t.run([&] { zero(); k_tma_dynamic<2><<<sms, 64, ENV_BYTES>>>(buf, nenv, parts, 8, counter, nullptr); }));
    snprintf(nm, sizeof nm, "K4d TMA DYNAMIC 4 issuers/SM, 8 env/grab, split %d", parts);
    rep(nm, t.run([&] { zero(); k_tma_dynamic<4><<<sms, 128, ENV_BYTES>>>(buf, nenv, parts, 8, counter, nullptr); }));
  }
  // straggler check: per-CTA finish time spread, static vs dynamic
  {
    std::vector<unsigned long long> f(sms);
    for (int dyn = 0; dyn < 2; ++dyn) {
      zero();
      CK(cudaDeviceSynchronize());
      if (dyn) k_tma_dynamic<1><<<sms, 32, ENV_BYTES>>>(buf, nenv, 1, 8, counter, finish);
      else k_tma_static_timed<<<sms, 32, ENV_BYTES>>>(buf, nenv, finish);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(f.data(), finish, 8 * sms, cudaMemcpyDeviceToHost));
      unsigned long long mn = ~0ull, mx = 0;
      for (auto v : f) { if (v < mn) mn = v; if (v > mx) mx = v; }
      std::vector<double> rel;
      for (auto v : f) rel.push_back((v - mn) / 1e6);
      std::sort(rel.begin(), rel.end());
      printf("finish-time spread (%s): first CTA done at 0, median +%.3f ms, p90 +%.3f ms, last +%.3f ms\n",
             dyn ? "dynamic" : "static ", rel[sms / 2], rel[sms * 9 / 10], rel[sms - 1]);
    }
  }
  CK(cudaFree(counter));
  CK(cudaFree(finish));
  CK(cudaFree(buf));
  return 0;
}
