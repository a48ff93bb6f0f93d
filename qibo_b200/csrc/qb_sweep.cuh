// qibo_b200 K2: the multi-gate sweep kernel (sm_100a).
//
// One persistent CTA per SM.  Warp 0 (LOADER) and warp 1 (STORER) move tiles between HBM and shared
// memory with bulk-async copies (cp.async.bulk, the TMA engine; SASS UBLKCP) -- one copy per contiguous
// run of the tile -- tracked by mbarriers (loads: complete_tx; stores: bulk groups).  Warps 2..9 are
// COMPUTE warps: they apply every gate of the sweep to the resident tile in shared memory.  Three 64 KiB tiles rotate
// through the states LOADING -> COMPUTING -> STORING, so HBM reads, shared-memory math and HBM writes of
// three consecutive tiles overlap.  HBM traffic per sweep: 2 * B * 2^n regardless of how many gates ride.
//
// The device program (SweepHeader + DevOp[] + payloads, qb_planner.hpp) is copied to shared memory once.
#pragma once
#include <cuda_runtime.h>

#include "qb_common.cuh"
#include "qb_gate_kernels.cuh"
#include "qb_planner.hpp"
#include "qb_passes.cuh"

namespace qb {

constexpr int SW_NBUF = 3;
constexpr int SW_TILE_BYTES = 1 << SWEEP_TILE_BYTES_LOG2;
#ifndef QB_COMPUTE_THREADS
#define QB_COMPUTE_THREADS 256  // 8 warps x 2 groups in flight per thread = the 512 groups of a 64 KiB tile
#endif
constexpr int SW_COMPUTE_THREADS = QB_COMPUTE_THREADS;
constexpr int SW_COPY_THREADS = 64;  // warp 0: loader, warp 1: storer
constexpr int SW_THREADS = SW_COMPUTE_THREADS + SW_COPY_THREADS;
constexpr int SW_BLOB_REGION = 30 * 1024;
constexpr int SW_MAX_RUNS = 256;
constexpr int SW_SMEM_BYTES = SW_NBUF * SW_TILE_BYTES + SW_BLOB_REGION + SW_MAX_RUNS * 8 + SWEEP_MAX_SLOTS * (16 + 4 + 4) + 128;
static_assert(SW_SMEM_BYTES <= 227 * 1024, "sweep kernel shared memory exceeds 227 KB");

// ---- PTX wrappers -------------------------------------------------------------------------------------
QB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
QB_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
QB_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
QB_D void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
QB_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (error returned to the host) instead of hanging the GPU
QB_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) __trap();
  }
}
// for the copy warps: they wait for whole tile periods, so back off between polls instead of competing with the
// compute warps of their scheduler for issue slots
QB_D void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(1000);
    if (clock64() - t0 > 20000000000LL) __trap();
  }
}
QB_D void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
QB_D void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
QB_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> QB_D void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
QB_D void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
QB_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
QB_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
QB_D void compute_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SW_COMPUTE_THREADS) : "memory"); }

// ---- the kernel -----------------------------------------------------------------------------------------
template <typename C>
__global__ void __launch_bounds__(SW_THREADS, 1) sweep_kernel(C* __restrict__ state, const char* __restrict__ prog) {
  extern __shared__ __align__(1024) unsigned char smem[];
  C* tiles = reinterpret_cast<C*>(smem);
  char* blob = reinterpret_cast<char*>(smem + SW_NBUF * SW_TILE_BYTES);
  uint64_t* run_off = reinterpret_cast<uint64_t*>(blob + SW_BLOB_REGION);
  double2* op_scal_raw = reinterpret_cast<double2*>(run_off + SW_MAX_RUNS);
  uint32_t* op_flag = reinterpret_cast<uint32_t*>(op_scal_raw + SWEEP_MAX_SLOTS);
  uint32_t* op_aux = op_flag + SWEEP_MAX_SLOTS;
  uint64_t* full = reinterpret_cast<uint64_t*>(op_aux + SWEEP_MAX_SLOTS);
  uint64_t* done = full + SW_NBUF;
  uint64_t* freeb = done + SW_NBUF;
  C* op_scal = reinterpret_cast<C*>(op_scal_raw);

  const int tid = threadIdx.x;
  {
    const uint32_t nbytes = reinterpret_cast<const SweepHeader*>(prog)->blob_bytes;
    for (uint32_t i = tid * 16; i < nbytes; i += SW_THREADS * 16)
      *reinterpret_cast<int4*>(blob + i) = __ldg(reinterpret_cast<const int4*>(prog + i));
  }
  __syncthreads();
  const SweepHeader& hdr = *reinterpret_cast<const SweepHeader*>(blob);
  const int T = (int)hdr.T, L = (int)hdr.L;
  const uint32_t nruns = 1u << (T - L);
  const uint32_t run_bytes = (uint32_t)sizeof(C) << L;
  const uint32_t tile_bytes = (uint32_t)sizeof(C) << T;
  const uint32_t tile_elems = 1u << T;
  for (uint32_t r = tid; r < nruns; r += SW_THREADS) run_off[r] = deposit(uint64_t(r) << L, hdr.tile_mask);
  if (tid == 0) {
    for (int b = 0; b < SW_NBUF; ++b) {
      mbar_init(&full[b], 1);
      mbar_init(&done[b], 1);
      mbar_init(&freeb[b], 1);
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();

  const uint64_t ntiles = hdr.ntiles;
  const uint64_t my_n = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const PassHeader* passes = reinterpret_cast<const PassHeader*>(blob + hdr.passes_offset);
  const uint32_t* slot_table = reinterpret_cast<const uint32_t*>(blob + hdr.slots_offset);
  const int npasses = (int)hdr.npasses;
  const int nslots = (int)hdr.nslots;

  if (tid < 32) {
    // ================= loader warp: HBM -> shared =================
    const int lane = tid;
    for (uint64_t i = 0; i < my_n; ++i) {
      const int b = (int)(i % SW_NBUF);
      // buffer b was last used by tile i-3: wait until the storer has drained it
      if (i >= SW_NBUF) mbar_wait_sleep(&freeb[b], (uint32_t)(((i / SW_NBUF) - 1) & 1));
      if (lane == 0) mbar_expect_tx(&full[b], tile_bytes);
      __syncwarp();
      const C* gbase = state + deposit(blockIdx.x + i * gridDim.x, hdr.other_mask);
      C* sbase = tiles + (size_t)b * tile_elems;
      for (uint32_t r = lane; r < nruns; r += 32) bulk_g2s(sbase + ((size_t)r << L), gbase + run_off[r], run_bytes, &full[b]);
    }
  } else if (tid < 64) {
    // ================= storer warp: shared -> HBM =================
    const int lane = tid - 32;
    for (uint64_t j = 0; j < my_n; ++j) {
      const int b = (int)(j % SW_NBUF);
      mbar_wait_sleep(&done[b], (uint32_t)((j / SW_NBUF) & 1));
      C* gbase = state + deposit(blockIdx.x + j * gridDim.x, hdr.other_mask);
      const C* sbase = tiles + (size_t)b * tile_elems;
      for (uint32_t r = lane; r < nruns; r += 32) bulk_s2g(gbase + run_off[r], sbase + ((size_t)r << L), run_bytes);
      bulk_commit();
      bulk_wait_read<0>();  // shared memory of this tile has been read out: the loader may refill it
      __syncwarp();
      if (lane == 0) mbar_arrive(&freeb[b]);
    }
    bulk_wait_all();
  } else {
    // ================= compute warps =================
    const uint32_t ctid = tid - SW_COPY_THREADS;
    for (uint64_t i = 0; i < my_n; ++i) {
      const int b = (int)(i % SW_NBUF);
      C* tile = tiles + (size_t)b * tile_elems;
      const uint64_t base = deposit(blockIdx.x + i * gridDim.x, hdr.other_mask);
      for (int sl = (int)ctid; sl < nslots; sl += SW_COMPUTE_THREADS) {
        const uint32_t so = slot_table[sl];
        if (so & 0x80000000u) {
          const DevOp& bop = *reinterpret_cast<const DevOp*>(blob + (so & 0x7fffffffu));
          op_flag[sl] = (base & bop.ext_cmask) == bop.ext_cmask ? 1u : 0u;
        } else {
          micro_prephase<C>(*reinterpret_cast<MicroOp*>(blob + so), blob, base, T);
        }
      }
      mbar_wait(&full[b], (uint32_t)((i / SW_NBUF) & 1));
      compute_bar();
      for (int pi = 0; pi < npasses; ++pi) {
        const PassHeader& ph = passes[pi];
        if (ph.kind == PASS_REGTILE) {
          switch (ph.R) {
            // R < 3 only occurs for n < 3, which qb_apply_program routes to the K1 kernels
            case 3: run_regtile<C, 3>(tile, blob, ph, T, ctid, SW_COMPUTE_THREADS); break;
            default: run_regtile<C, 4>(tile, blob, ph, T, ctid, SW_COMPUTE_THREADS); break;
          }
        } else {
          const DevOp& op = *reinterpret_cast<const DevOp*>(blob + ph.offset);
          if (op_flag[op.slot]) {
            const C* payload = op.payload_global ? reinterpret_cast<const C*>(prog + op.payload) : reinterpret_cast<const C*>(blob + op.payload);
            const uint32_t ntasks = (1u << (T - (int)op.nins)) << (op.k - 3);
            for (uint32_t t = 0; t < ntasks; t += SW_COMPUTE_THREADS) {
              BigAcc<C> a;
              big_read<C>(tile, op, payload, T, t + ctid, a);
              compute_bar();
              big_write<C>(tile, op, a);
              if (t + SW_COMPUTE_THREADS < ntasks) compute_bar();
            }
          }
        }
        compute_bar();
      }
      fence_proxy_async();
      compute_bar();
      if (ctid == 0) mbar_arrive(&done[b]);
    }
  }
}

inline int sweep_configure(const cudaDeviceProp& prop) {
  if ((int)prop.sharedMemPerBlockOptin < SW_SMEM_BYTES) return QB_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(sweep_kernel<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES) != cudaSuccess)
    return QB_ERR_CUDA;
  if (cudaFuncSetAttribute(sweep_kernel<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES) != cudaSuccess)
    return QB_ERR_CUDA;
  return QB_OK;
}

inline int launch_sweep(cudaStream_t stream, int sm_count, void* state, int nqubits, int dtype, const SweepDesc& sd,
                        const char* prog_dev) {
  (void)nqubits;
  uint64_t grid = sd.ntiles < (uint64_t)sm_count ? sd.ntiles : (uint64_t)sm_count;
  if (dtype == QB_C128)
    sweep_kernel<double2><<<(unsigned)grid, SW_THREADS, SW_SMEM_BYTES, stream>>>((double2*)state, prog_dev + sd.blob_offset);
  else
    sweep_kernel<float2><<<(unsigned)grid, SW_THREADS, SW_SMEM_BYTES, stream>>>((float2*)state, prog_dev + sd.blob_offset);
  return cudaPeekAtLastError() == cudaSuccess ? QB_OK : QB_ERR_CUDA;
}

// ---- K7 helper: gather / scatter the half of the shard with bit `pos` == `bit` ----------------------------
template <typename C>
__global__ void __launch_bounds__(256) k7_half_copy(C* __restrict__ state, C* __restrict__ staging, uint64_t half, int pos, int bit,
                                                    int unpack) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t g = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < half; g += stride) {
    const uint64_t idx = insert_zero(g, pos) | (uint64_t(bit) << pos);
    if (unpack) st_stream(state + idx, ld_stream(staging + g));
    else st_stream(staging + g, ld_stream(state + idx));
  }
}

// ---- K7: pairwise half-shard swap through NVLink peer memory ------------------------------------------------
// mine[a] <-> peer[b] for every index i of this rank's share; 4 independent pairs in flight per thread (remote
// latency is ~2 us: bytes in flight, not arithmetic, set the rate).
template <typename C, int U>
__global__ void __launch_bounds__(256) k7_swap_half_p2p(C* __restrict__ mine, C* __restrict__ peer, int pos, int mybit, uint64_t begin,
                                                        uint64_t end) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  const uint64_t ma = uint64_t(1 - mybit) << pos, mb = uint64_t(mybit) << pos;
  for (uint64_t i0 = begin + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i0 < end; i0 += stride * U) {
    C x[U], y[U];
    uint64_t base[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + uint64_t(u) * stride;
      base[u] = insert_zero(i, pos);
      if (i < end) {
        x[u] = ld_stream(mine + (base[u] | ma));
        y[u] = ld_stream(peer + (base[u] | mb));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + uint64_t(u) * stride;
      if (i < end) {
        st_stream(mine + (base[u] | ma), y[u]);
        st_stream(peer + (base[u] | mb), x[u]);
      }
    }
  }
}

inline int launch_swap_half_p2p(cudaStream_t stream, int sm_count, void* state, void* peer, int nqubits, int dtype, int pos, int mybit,
                                int part, int nparts) {
  const uint64_t half = uint64_t(1) << (nqubits - 1);
  const uint64_t begin = half / nparts * part, end = part == nparts - 1 ? half : half / nparts * (part + 1);
  const int grid = sm_count * env_int("QB_P2P_BLOCKS_PER_SM", 8);
  const int unroll = env_int("QB_P2P_UNROLL", 4);
  if (dtype == QB_C128) {
    if (unroll >= 8) k7_swap_half_p2p<double2, 8><<<grid, 256, 0, stream>>>((double2*)state, (double2*)peer, pos, mybit, begin, end);
    else if (unroll >= 4) k7_swap_half_p2p<double2, 4><<<grid, 256, 0, stream>>>((double2*)state, (double2*)peer, pos, mybit, begin, end);
    else k7_swap_half_p2p<double2, 2><<<grid, 256, 0, stream>>>((double2*)state, (double2*)peer, pos, mybit, begin, end);
  } else {
    if (unroll >= 8) k7_swap_half_p2p<float2, 8><<<grid, 256, 0, stream>>>((float2*)state, (float2*)peer, pos, mybit, begin, end);
    else k7_swap_half_p2p<float2, 4><<<grid, 256, 0, stream>>>((float2*)state, (float2*)peer, pos, mybit, begin, end);
  }
  return cudaPeekAtLastError() == cudaSuccess ? QB_OK : QB_ERR_CUDA;
}

inline int launch_half_copy(cudaStream_t stream, int sm_count, void* state, void* staging, int nqubits, int dtype, int pos, int bit,
                            int unpack) {
  const uint64_t half = uint64_t(1) << (nqubits - 1);
  uint64_t blocks = (half + 255) / 256;
  uint64_t cap = (uint64_t)sm_count * 32;
  int grid = (int)(blocks < cap ? blocks : cap);
  if (dtype == QB_C128)
    k7_half_copy<double2><<<grid, 256, 0, stream>>>((double2*)state, (double2*)staging, half, pos, bit, unpack);
  else
    k7_half_copy<float2><<<grid, 256, 0, stream>>>((float2*)state, (float2*)staging, half, pos, bit, unpack);
  return cudaPeekAtLastError() == cudaSuccess ? QB_OK : QB_ERR_CUDA;
}
}  // namespace qb
