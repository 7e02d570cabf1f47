// Per-CTA top-k inside the scoring epilogue: replaces the M-float acquisition vector + the radix select over
// it for `jnp.argsort(acq_vals)[-k:]` (optimizer.py:196).
//
// Keys are (orderable(value) << 32 | global index): unique, ordered like a stable ascending argsort (NaN largest, higher
// index wins among equal values), so "the k largest keys" is the reference's index set even when millions of values tie.
// Every CTA keeps a threshold key (the k-th largest it has seen) in shared memory and appends candidates above it to
// its own slab of a global buffer; a random stream of m candidates appends ~k ln(m / k) of them.  When the slab is nearly
// full the 128 epilogue threads prune it: exact radix select of the k-th largest key (8 passes over keys held in
// registers, a 256-bin shared-memory histogram), survivors re-appended, threshold raised.  The union of the CTAs'
// survivors contains the global top-k whatever the tile schedule, so the result is deterministic; topk_merge_kernel
// (topk.cu) reduces grid * k survivors to the final sorted list.
#pragma once
#include "common.cuh"

namespace bx {
namespace tk {

constexpr int CAP = 2048;        // keys per CTA slab (>= BX_MAX_K + 2 * TPB)
constexpr int TPB = 128;         // threads that cooperate (the four epilogue warps)
constexpr int KPT = CAP / TPB;   // keys per thread while pruning

struct Shared {
  unsigned long long thr;   // insert iff key > thr (0: accept everything)
  unsigned long long prefix;
  unsigned int cnt, rem, hist[256];
};

__device__ __forceinline__ unsigned long long topk_key(float v, uint32_t index) {
  return ((unsigned long long)orderable_u32(v) << 32) | index;
}

template <int BAR_ID>
__device__ __forceinline__ void bar() { asm volatile("bar.sync %0, %1;" ::"n"(BAR_ID), "n"(TPB) : "memory"); }

__device__ __forceinline__ void init(Shared& S, int t) {
  if (t == 0) { S.thr = 0ull; S.cnt = 0u; }
}

__device__ __forceinline__ void push(Shared& S, unsigned long long* slab, unsigned long long key) {
  if (key > S.thr) {
    const unsigned int slot = atomicAdd(&S.cnt, 1u);
    if (slot < (unsigned)CAP) slab[slot] = key;  // never false: prune() runs whenever fewer than TPB slots are left
  }
}

// All TPB threads.  Keeps the k largest keys of the slab (no-op when it holds at most k) and raises the threshold.
template <int BAR_ID>
__device__ __noinline__ void prune(Shared& S, unsigned long long* slab, int k, int t) {
  bar<BAR_ID>();
  const unsigned int cnt = min(S.cnt, (unsigned)CAP);
  if (cnt <= (unsigned)k) return;  // uniform
  unsigned long long mine[KPT];
#pragma unroll
  for (int i = 0; i < KPT; ++i) {
    const unsigned int e = (unsigned)(i * TPB + t);
    mine[i] = e < cnt ? slab[e] : 0ull;  // 0 is no key (orderable(-inf) > 0)
  }
  if (t == 0) { S.prefix = 0ull; S.rem = (unsigned)k; }
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    S.hist[t] = 0u;
    S.hist[t + TPB] = 0u;
    bar<BAR_ID>();
    const unsigned long long prefix = S.prefix;
#pragma unroll
    for (int i = 0; i < KPT; ++i)
      if (mine[i] != 0ull && (pass == 0 || (mine[i] >> (shift + 8)) == prefix)) atomicAdd(&S.hist[(mine[i] >> shift) & 0xFF], 1u);
    bar<BAR_ID>();
    if (t < 32) {  // warp 0: bin of the rem-th largest key among those that share the prefix
      unsigned int c[8], s = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { c[i] = S.hist[t * 8 + i]; s += c[i]; }
      unsigned int incl = s;  // inclusive suffix sum over the lanes >= t (higher lanes hold the larger bins)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_down_sync(0xffffffffu, incl, o);
        if (t + o < 32) incl += v;
      }
      const unsigned int above = incl - s;  // keys in the bins of higher lanes
      const unsigned int rem = S.rem;
      if (above < rem && rem <= above + s) {
        unsigned int r = rem - above;
        int b = 7;
        for (; b > 0; --b) {
          if (c[b] >= r) break;
          r -= c[b];
        }
        S.prefix = (prefix << 8) | (unsigned long long)(t * 8 + b);
        S.rem = r;
      }
    }
    bar<BAR_ID>();
  }
  const unsigned long long thr = S.prefix;  // the k-th largest key itself
  bar<BAR_ID>();
  if (t == 0) { S.cnt = 0u; S.thr = thr; }
  bar<BAR_ID>();
#pragma unroll
  for (int i = 0; i < KPT; ++i)
    if (mine[i] >= thr && mine[i] != 0ull) slab[atomicAdd(&S.cnt, 1u)] = mine[i];
  bar<BAR_ID>();
}

// Call once per tile after the pushes (all TPB threads): prunes when the next tile could overflow the slab.
template <int BAR_ID>
__device__ __forceinline__ void end_tile(Shared& S, unsigned long long* slab, int k, int t) {
  bar<BAR_ID>();
  const unsigned int cnt = S.cnt;
  bar<BAR_ID>();  // nobody pushes for the next tile before everyone has read the count: the decision is uniform
  if (cnt > (unsigned)(CAP - TPB)) prune<BAR_ID>(S, slab, k, t);
}

// Call once at the end of the kernel: at most k keys stay in the slab, their count goes to *count_out.
template <int BAR_ID>
__device__ __forceinline__ void finish(Shared& S, unsigned long long* slab, int k, int t, unsigned int* count_out) {
  prune<BAR_ID>(S, slab, k, t);
  bar<BAR_ID>();
  if (t == 0) *count_out = min(S.cnt, (unsigned)CAP);
}

}  // namespace tk
}  // namespace bx
