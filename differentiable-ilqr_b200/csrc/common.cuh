// common.cuh -- shared device helpers for the dilqr sm_100a kernels.
//
// Build note: every translation unit is compiled with -fmad=false and all
// contractions are written as explicit fma() calls.  The line search compares
// the cost of a new trajectory with the cost recorded for the current one
// (lqr_step.py:169,176-179); the two values are produced by different kernels /
// inlining contexts, and they must be bit-identical when the trajectories are
// identical (fixed point).  Disabling implicit contraction makes the rounding
// sequence a property of the source, not of the surrounding code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace dilqr {

#define DILQR_DEVICE __device__ __forceinline__

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kPnqpMaxIter = 20;   // pnqp.py:5
constexpr int kArmijoMax = 10;     // pnqp.py:65

template <class S> DILQR_DEVICE S fmaS(S a, S b, S c);
template <> DILQR_DEVICE float fmaS<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> DILQR_DEVICE double fmaS<double>(double a, double b, double c) { return fma(a, b, c); }

template <class S> DILQR_DEVICE S sqrtS(S a);
template <> DILQR_DEVICE float sqrtS<float>(float a) { return sqrtf(a); }
template <> DILQR_DEVICE double sqrtS<double>(double a) { return sqrt(a); }

template <class S> DILQR_DEVICE S absS(S a);
template <> DILQR_DEVICE float absS<float>(float a) { return fabsf(a); }
template <> DILQR_DEVICE double absS<double>(double a) { return fabs(a); }

template <class S> DILQR_DEVICE S expS(S a);
template <> DILQR_DEVICE float expS<float>(float a) { return expf(a); }
template <> DILQR_DEVICE double expS<double>(double a) { return exp(a); }
template <class S> DILQR_DEVICE S atan2S(S y, S x);
template <> DILQR_DEVICE float atan2S<float>(float y, float x) { return atan2f(y, x); }
template <> DILQR_DEVICE double atan2S<double>(double y, double x) { return atan2(y, x); }

template <class S> DILQR_DEVICE void sincosS(S a, S* s, S* c);
template <> DILQR_DEVICE void sincosS<float>(float a, float* s, float* c) { sincosf(a, s, c); }
template <> DILQR_DEVICE void sincosS<double>(double a, double* s, double* c) { sincos(a, s, c); }

// compile-time loop: f(std::integral_constant<int, I>{}) for I in [0, N) -- used where an
// index must be a constant expression (packed-table offsets computed by constexpr code)
template <int I, int N, class F>
DILQR_DEVICE void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// eclamp (util.py:58-72): two masked assignments, NaN passes through.
template <class S> DILQR_DEVICE S eclamp(S x, S lo, S hi) {
  if (x < lo) x = lo;
  if (x > hi) x = hi;
  return x;
}

// ---------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA, SASS: UBLKCP) helpers
// ---------------------------------------------------------------------------
DILQR_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
DILQR_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
DILQR_DEVICE void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
DILQR_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
DILQR_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
DILQR_DEVICE void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Software prefetch of the cache line holding *p (kernels that read AoS operands directly).
DILQR_DEVICE void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// ---------------------------------------------------------------------------
// Per-warp double-buffered slab stager.
//
// The API tensors are time-major AoS: for a fixed t the data of the 32
// consecutive problems a warp owns is ONE contiguous slab
// (C[t, b0:b0+32] = 32*n*n scalars).  One elected lane moves each slab into
// shared memory with a single cp.async.bulk (TMA) tracked by an mbarrier; every
// lane then reads its own problem's block.  Two stages per warp, so the slab of
// the next timestep is in flight while the current one is consumed.  Warps never
// synchronise with each other.  Partial / unaligned slabs (tail warp) fall back
// to a lane-strided copy.
// ---------------------------------------------------------------------------
constexpr int kStages = 2;

// rarely taken path (tail warps / unaligned slabs): kept out of line so the hot
// sweeps stay small
template <class S>
__device__ __noinline__ void lane_copy(S* dst, const S* src, int cnt, int lane) {
  for (int e = lane; e < cnt; e += kWarp) dst[e] = __ldg(src + e);
}

template <class S, int kMaxSeg = 6>
struct WarpStager {
  char* base;        // this warp's staging area (kStages * stage_bytes)
  uint64_t* bar;     // kStages barriers
  uint32_t stage_bytes;
  uint32_t seg_off[kMaxSeg];    // byte offset of each segment in a stage
  uint32_t seg_elems[kMaxSeg];  // scalars per problem in each segment
  uint32_t seg_nbytes[kMaxSeg]; // bytes copied per issue for each segment
  uint32_t seg_sized;           // bit i: seg_nbytes[i] is a non-zero multiple of 16
  uint32_t seg_full;            // bit i: segment i always holds 32 problems (workspace chunks)
  uint32_t seg_shared;          // bit i: segment i is ONE block shared by all lanes
                                //        (broadcast cost: C[n,n] / C[T,n,n], mpc.py:205-219)
  uint32_t parity;   // bit s = parity to wait for on stage s
  uint32_t via_tma;  // bit s = stage s has bulk copies in flight
  uint32_t via_lane; // bit s = stage s has lane-copied segments
  int nvalid;        // problems this warp really has (<= 32)
  int lane;
  // bound sources (bind / issue_bound): segment i of timestep t lives at
  // bnd_base[i] + t * bnd_stride[i] (bytes); bnd_fast bit i = every such address and
  // the copy size are 16-byte multiples, so the segment can always go by bulk copy
  const char* bnd_base[kMaxSeg];
  long long bnd_stride[kMaxSeg];
  uint32_t bnd_fast;

  DILQR_DEVICE uint32_t seg_bytes(int i) const {
    const uint32_t cnt = ((seg_shared >> i) & 1u) ? 1u : (((seg_full >> i) & 1u) ? kWarp : nvalid);
    return seg_elems[i] * cnt * (uint32_t)sizeof(S);
  }

  // `full_mask` bit i = segment i is a warp-blocked workspace chunk [elems][32]
  // (always complete, lane-interleaved) rather than an API slab [nvalid][elems];
  // `shared_mask` bit i = a single [elems] block read by every lane.
  DILQR_DEVICE void init(char* smem_base, uint64_t* bars, int lane_, int nvalid_, int nseg,
                         const uint32_t* elems, uint32_t full_mask = 0, uint32_t shared_mask = 0) {
    base = smem_base;
    seg_full = full_mask;
    seg_shared = shared_mask;
    bar = bars;
    lane = lane_;
    nvalid = nvalid_;
    parity = 0;
    via_tma = 0;
    via_lane = 0;
    bnd_fast = 0;
#pragma unroll
    for (int i = 0; i < kMaxSeg; ++i) {
      bnd_base[i] = nullptr;
      bnd_stride[i] = 0;
    }
    uint32_t off = 0;
    seg_sized = 0;
#pragma unroll
    for (int i = 0; i < kMaxSeg; ++i) {
      seg_off[i] = off;
      seg_elems[i] = i < nseg ? elems[i] : 0;
      seg_nbytes[i] = seg_bytes(i);
      if (seg_nbytes[i] && !(seg_nbytes[i] & 15u)) seg_sized |= 1u << i;
      uint32_t full = seg_elems[i] * kWarp * sizeof(S);
      off += (full + 15u) & ~15u;
    }
    stage_bytes = off;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < kStages; ++s) mbar_init(&bar[s], 1);
      mbar_fence_init();
    }
    __syncwarp();
  }

  // New segment table for the next sweep of the same kernel (all copies of the previous
  // sweep have been waited for): offsets / sizes are recomputed, barriers and their
  // parity are kept, bound sources are cleared.
  DILQR_DEVICE void reconfigure(int nseg, const uint32_t* elems, uint32_t full_mask = 0,
                                uint32_t shared_mask = 0) {
    __syncwarp();
    seg_full = full_mask;
    seg_shared = shared_mask;
    bnd_fast = 0;
    uint32_t off = 0;
    seg_sized = 0;
#pragma unroll
    for (int i = 0; i < kMaxSeg; ++i) {
      bnd_base[i] = nullptr;
      bnd_stride[i] = 0;
      seg_off[i] = off;
      seg_elems[i] = i < nseg ? elems[i] : 0;
      seg_nbytes[i] = seg_bytes(i);
      if (seg_nbytes[i] && !(seg_nbytes[i] & 15u)) seg_sized |= 1u << i;
      uint32_t full = seg_elems[i] * kWarp * sizeof(S);
      off += (full + 15u) & ~15u;
    }
    stage_bytes = off;
  }

  // change which segments are shared blocks (a stage reused for different tensors)
  DILQR_DEVICE void set_shared(uint32_t shared_mask) {
    seg_shared = shared_mask;
    seg_sized = 0;
#pragma unroll
    for (int i = 0; i < kMaxSeg; ++i) {
      seg_nbytes[i] = seg_bytes(i);
      if (seg_nbytes[i] && !(seg_nbytes[i] & 15u)) seg_sized |= 1u << i;
    }
  }

  static __host__ __device__ uint32_t bytes_per_warp(int nseg, const uint32_t* elems) {
    uint32_t off = 0;
    for (int i = 0; i < nseg; ++i) off += (elems[i] * kWarp * (uint32_t)sizeof(S) + 15u) & ~15u;
    return off * kStages;
  }

  // Start copying the operands of one timestep into `stage`.  src[i] points at the
  // first scalar of this warp's slab of segment i (or nullptr to skip).  Fast path
  // (every requested segment 16-byte sized and aligned -- full warps, even shapes):
  // one elected lane posts the byte count and one bulk TMA copy per segment.  Slow
  // path (tail warps, odd sizes): per segment, bulk if possible else a lane copy.
  DILQR_DEVICE void issue(int stage, const S* const* src, int nseg) {
    char* dst = base + stage * stage_bytes;
    __syncwarp();  // all lanes are done reading this stage (WAR)
    uintptr_t orp = 0;
    uint32_t need = 0;
#pragma unroll
    for (int i = 0; i < kMaxSeg; ++i) {
      if (i < nseg && src[i]) {
        orp |= reinterpret_cast<uintptr_t>(src[i]);
        need |= 1u << i;
      }
    }
    if (((need & ~seg_sized) == 0u) && !(orp & 15u)) {
      via_tma |= 1u << stage;
      via_lane &= ~(1u << stage);
      if (lane == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int i = 0; i < kMaxSeg; ++i)
          if (i < nseg && src[i]) total += seg_nbytes[i];
        mbar_expect_tx(&bar[stage], total);
#pragma unroll
        for (int i = 0; i < kMaxSeg; ++i)
          if (i < nseg && src[i]) bulk_g2s(dst + seg_off[i], src[i], seg_nbytes[i], &bar[stage]);
      }
      return;
    }
    uint32_t bulk_mask = 0, total = 0;
#pragma unroll
    for (int i = 0; i < kMaxSeg; ++i) {
      if (i < nseg && src[i] && ((seg_sized >> i) & 1u) &&
          !(reinterpret_cast<uintptr_t>(src[i]) & 15u)) {
        bulk_mask |= 1u << i;
        total += seg_nbytes[i];
      }
    }
    via_tma = bulk_mask ? (via_tma | (1u << stage)) : (via_tma & ~(1u << stage));
    via_lane |= 1u << stage;
    if (bulk_mask && lane == 0) {
      mbar_expect_tx(&bar[stage], total);
#pragma unroll
      for (int i = 0; i < kMaxSeg; ++i)
        if ((bulk_mask >> i) & 1u) bulk_g2s(dst + seg_off[i], src[i], seg_nbytes[i], &bar[stage]);
    }
#pragma unroll
    for (int i = 0; i < kMaxSeg; ++i) {
      if (i < nseg && src[i] && !((bulk_mask >> i) & 1u))
        lane_copy<S>(reinterpret_cast<S*>(dst + seg_off[i]), src[i], seg_nbytes[i] / sizeof(S), lane);
    }
  }

  // Describe where segment i comes from for the rest of a sweep (nullptr: never issued).
  DILQR_DEVICE void bind(int i, const void* base, long long stride_bytes) {
    bnd_base[i] = static_cast<const char*>(base);
    bnd_stride[i] = stride_bytes;
    const bool ok = base && ((seg_sized >> i) & 1u) &&
                    !(reinterpret_cast<uintptr_t>(base) & 15u) && !(stride_bytes & 15);
    bnd_fast = ok ? (bnd_fast | (1u << i)) : (bnd_fast & ~(1u << i));
  }

  // issue() for bound sources: `mask` = segments wanted for timestep t.  When all of
  // them are bulk-eligible only the elected lane computes addresses.
  DILQR_DEVICE void issue_bound(int stage, int t, uint32_t mask) {
    if ((mask & ~bnd_fast) == 0u) {
      __syncwarp();  // all lanes are done reading this stage (WAR)
      via_tma |= 1u << stage;
      via_lane &= ~(1u << stage);
      if (lane == 0) {
        char* dst = base + stage * stage_bytes;
        uint32_t total = 0;
#pragma unroll
        for (int i = 0; i < kMaxSeg; ++i)
          if ((mask >> i) & 1u) total += seg_nbytes[i];
        mbar_expect_tx(&bar[stage], total);
#pragma unroll
        for (int i = 0; i < kMaxSeg; ++i)
          if ((mask >> i) & 1u)
            bulk_g2s(dst + seg_off[i], bnd_base[i] + (long long)t * bnd_stride[i], seg_nbytes[i],
                     &bar[stage]);
      }
      return;
    }
    const S* src[kMaxSeg];
#pragma unroll
    for (int i = 0; i < kMaxSeg; ++i)
      src[i] = (((mask >> i) & 1u) && bnd_base[i])
                   ? reinterpret_cast<const S*>(bnd_base[i] + (long long)t * bnd_stride[i])
                   : nullptr;
    issue(stage, src, kMaxSeg);
  }

  DILQR_DEVICE void wait(int stage) {
    if ((via_tma >> stage) & 1u) {
      mbar_wait(&bar[stage], (parity >> stage) & 1u);
      parity ^= (1u << stage);
    }
    if ((via_lane >> stage) & 1u) __syncwarp();   // lane-copied segments
  }

  // Base of segment i in `stage` (blocked chunks: element e of this lane is [e*32+lane]).
  DILQR_DEVICE const S* seg_ptr(int stage, int seg) const {
    return reinterpret_cast<const S*>(base + stage * stage_bytes + seg_off[seg]);
  }

  // Pointer to this lane's block of segment i in `stage`.
  DILQR_DEVICE const S* lane_ptr(int stage, int seg) const {
    return reinterpret_cast<const S*>(base + stage * stage_bytes + seg_off[seg]) +
           (((seg_shared >> seg) & 1u) ? 0 : lane * seg_elems[seg]);
  }
};

// warp-blocked SoA index of workspace arrays [T][B/32][ncomp][32]
DILQR_DEVICE size_t bidx(int t, int comp, int ncomp, int b, int nW) {
  return (((size_t)t * nW + (b >> 5)) * ncomp + comp) * kWarp + (b & 31);
}

// ordered-uint encoding of non-negative doubles for atomicMax
DILQR_DEVICE unsigned long long dbits(double v) { return (unsigned long long)__double_as_longlong(v); }

}  // namespace dilqr
