// K3/K4/K5/K9: destination-segmented gather-reduce for sm_100a (HBM-bound).
//
// Work decomposition
//   * a "lane group" of G = pow2ceil(F/VEC) <= 32 lanes covers one feature row; each lane keeps
//     NCH vectors (VEC floats, 128-bit loads when VEC = 4) of the row in registers;
//   * a task is a block of G consecutive CSR rows.  The group loads the G+1 row pointers once and
//     then WALKS THE BLOCK'S CONTIGUOUS EDGE RANGE: column indices (and weights) are fetched G at
//     a time with one coalesced access (the next batch is prefetched), broadcast with shuffles, and
//     U = 8/NCH feature-row loads per lane are in flight regardless of where row boundaries fall.
//     Row boundaries only decide when the accumulator is flushed through the epilogue, so short
//     rows (the common case in power-law graphs) no longer serialise three dependent memory
//     round trips each - the first version of this kernel did and sat at 19 % DRAM utilisation
//     (profiles/r01_gather_reduce_v1_raw.csv);
//   * edges are accumulated in CSR order, i.e. the order a sequential scatter visits them -
//     results are deterministic (no atomics) and, for unweighted sums of non-hub rows,
//     identical to the host reference;
//   * rows above `hub_threshold` edges are skipped by the walk; they are cut into chunks of
//     `hub_chunk` edges (scheduled first, they are the long tasks), reduced into a partial buffer
//     and merged in chunk order by hub_finish_kernel;
//   * persistent grid-stride over tasks, grid = min(tasks, SMs * resident CTAs).
#include <math.h>

#include "common.cuh"

namespace kgb {

// Destination table of a halo push (K7): slot s (grouped by destination peer) -> row of that peer's window.
struct PushTab {
  int n_peers;
  int64_t slot_begin[KGB_MAX_PEERS + 1];
  float* dst[KGB_MAX_PEERS];
  int64_t dst_row0[KGB_MAX_PEERS];
  int64_t ldd;
  // halo_push_kernel only: the walk visits 256-slot chunks in the order (k * perm_mul + perm_add) mod n_chunks, so at
  // any moment the stores of one rank are spread over all receivers in proportion to their blocks, and different
  // ranks start at different places - with the natural order all ranks store into peer 0 first, then all into
  // peer 1 ... and the transfer runs at one receiver's ingest rate (measured: 330 GB/s per sender at 8 GPUs)
  int64_t n_chunks, perm_mul, perm_add;
};
constexpr int PUSH_CHUNK = 256;

__device__ __forceinline__ float* push_row(const PushTab& tab, int64_t s) {
  int p = 0;
#pragma unroll
  for (int q = 1; q < KGB_MAX_PEERS; ++q) p += (q < tab.n_peers && s >= tab.slot_begin[q]) ? 1 : 0;
  return tab.dst[p] + (tab.dst_row0[p] + (s - tab.slot_begin[p])) * tab.ldd;
}

struct GRP {
  const float* x; int64_t ldx;
  int F;
  const int64_t* rowptr; const int32_t* col; int64_t n_rows;
  const int32_t* row_ids;
  const float* edge_w; const float* src_scale; const float* out_scale;
  const float* addend; int64_t ld_addend; float addend_scale;
  const float* bias; int act; int mean; int negate; int raw_max; int sqdev;
  float* out; int64_t ldo; int32_t* arg;
  const int32_t* hub_row; const int32_t* hub_chunk_base; const int32_t* hub_nchunks;
  const int32_t* chunk_hub;
  int n_hubs; int n_chunks; int hub_threshold; int hub_chunk;
  float* partial; int32_t* partial_arg;
  int32_t* work;  // [2] zero on entry: dynamic task queue head + finished-CTA count (self-resetting)
  const int32_t* unit_order;  // optional permutation of the row units (heaviest first)
  // split source / split output (partitioned graphs)
  const float* x2; int64_t ldx2; int64_t n_split_src;   // n_split_src = INT64_MAX when x2 is unused
  float* out2; int64_t ldo2; int64_t n_split_out;       // n_split_out = INT64_MAX when unused
  int has_push; PushTab tab;                            // fused halo push: split-output rows go to the peers' windows
  // optional copy of `col` whose bit 31 marks "hot" source rows (among the most frequently gathered rows that
  // together fit in the L2).  Hot rows are loaded with an L2 evict_last hint, all others with evict_first, so the
  // one-touch rows of a power-law graph stop pushing the hub rows out of the 126 MB L2.
  const int32_t* col_hot;
  // in-kernel dropout of the gathered rows (per edge and element, layers/gcn_conv.py:238-242, sage_conv.py:295-297):
  // edge_id[k] = original edge id of slot k, so the forward (CSR) and transposed (CSC) passes regenerate one mask
  const int32_t* edge_id; uint32_t drop_thr; float drop_scale; uint32_t seed_lo, seed_hi;
};

template <int VEC, int G, int NCH, bool IS_MAX>
__device__ __forceinline__ void init_acc(float (&acc)[NCH][VEC], int32_t (&aidx)[NCH][VEC]) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      acc[ch][e] = IS_MAX ? -INFINITY : 0.f;
      aidx[ch][e] = -1;
    }
}

// SQDEV (std aggregator, second pass): the mean of the row that is about to be reduced rides in the argmax slots
template <int VEC, int G, int NCH, bool SQDEV, class P>
__device__ __forceinline__ void load_mu(const P& p, int64_t row, int gl, const bool (&on)[NCH],
                                        int32_t (&aidx)[NCH][VEC]) {
  if constexpr (SQDEV) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float mu[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) mu[e] = 0.f;
      if (on[ch]) ld_vec<VEC>(p.addend + row * p.ld_addend + (gl + ch * G) * VEC, mu);
#pragma unroll
      for (int e = 0; e < VEC; ++e) aidx[ch][e] = __float_as_int(mu[e]);
    }
  }
}

__device__ __forceinline__ float fmax_nan(float a, float b) {
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));  // FMNMX.NAN: NaN propagates like torch's amax
  return d;
}

template <bool IS_MAX, bool HAS_W = true, bool SQDEV = false>
__device__ __forceinline__ void accum(float& m, int32_t& a, float v, float w, int32_t c, bool negate) {
  if constexpr (SQDEV) {
    // sum of squared deviations from the row's mean, whose bits ride in the (otherwise unused) argmax slot
    const float d = __fsub_rn(v, __int_as_float(a));
    m = __fadd_rn(m, __fmul_rn(d, d));   // square, then add: like segment_sum(square(m - mean)) (aggregators.py:211-216)
  } else if constexpr (IS_MAX) {
    // branch-free on purpose: an if/else chain compiles to divergent BSSY/BRA/BSYNC per element
    const float val = negate ? -v : v;
    const float mo = m;
    m = fmax_nan(mo, val);
    const bool gt = val > mo;
    const bool eq = val == mo;
    // tie between different sources: backward re-walks the row.  Parallel edges from the same source are
    // not a tie: their even shares add up to the whole gradient on that one row.
    const int32_t a_eq = (c != a) ? -2 : a;
    a = gt ? c : (eq ? a_eq : a);
  } else {
    if constexpr (HAS_W) m = __fadd_rn(m, __fmul_rn(w, v));  // mul then add, like message*w then segment_sum
    else m = __fadd_rn(m, v);
  }
}

// Load the feature rows of U consecutive slots of the current index batch.  Slots past the end of the
// batch carry index 0 (a valid row) and are simply never accumulated, so the loads need no predicate.
template <int VEC, int G, int NCH, int U, bool HAS_W, bool SPLIT, bool HOT, bool DROP>
__device__ __forceinline__ void load_batch_full(const GRP& p, int32_t myc, float myw, int32_t mye, int j, int gl,
                                                unsigned gmask, const bool (&on)[NCH], float (&v)[U][NCH][VEC],
                                                float (&w)[U], int32_t (&c)[U], uint32_t (&mk)[U][NCH]) {
  // Lanes beyond the row width re-read the row's first vector (same sector as lane 0, no extra traffic)
  // so every load is unconditional: a predicated load would make v loop-carried and spill.
  int loff[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) loff[ch] = on[ch] ? (gl + ch * G) * VEC : 0;
  uint64_t pol_last = 0, pol_cold = 0;
  if constexpr (HOT) {
    pol_last = l2_policy_evict_last();
    pol_cold = l2_policy_evict_first();
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    c[u] = __shfl_sync(gmask, myc, j + u, G);
    w[u] = 1.f;
    if constexpr (HAS_W) w[u] = __shfl_sync(gmask, myw, j + u, G);
    bool hot = false;
    if constexpr (HOT) {  // the index batch carries the hot flag in its sign bit (set when the batch was fetched)
      hot = c[u] < 0;
      c[u] &= 0x7fffffff;
    }
    const int64_t cu = c[u];
    const float* rp = p.x + cu * p.ldx;
    if constexpr (SPLIT) {
      if (cu >= p.n_split_src) rp = p.x2 + (cu - p.n_split_src) * p.ldx2;
    }
    if constexpr (HOT) {
      const uint64_t pol = hot ? pol_last : pol_cold;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) ld_vec_hint<VEC>(rp + loff[ch], v[u][ch], pol);
    } else {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) ld_vec<VEC>(rp + loff[ch], v[u][ch]);
    }
  }
  if constexpr (DROP) {
    // keep-bits of every loaded (edge, feature block): pure ALU work (Philox) issued behind the loads
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t eid = (uint32_t)__shfl_sync(gmask, mye, j + u, G);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const uint32_t f0 = (uint32_t)((gl + ch * G) * VEC);
        const uint32_t bits = dropout_keep4(eid, f0 >> 2, 0u, p.seed_lo, p.seed_hi, p.drop_thr);
        mk[u][ch] = (VEC == 4) ? bits : (bits >> (f0 & 3u));   // VEC < 4: this lane's elements start at f0 % 4
      }
    }
  }
}

// column index of a slot, with the hot flag of that source row in the sign bit (HOT only)
template <bool HOT>
__device__ __forceinline__ int32_t load_col(const GRP& p, int64_t k) {
  if constexpr (HOT) return __ldg(p.col_hot + k);   // same ids, hot flag already in bit 31
  return __ldg(p.col + k);
}

// Lanes whose `on` is false accumulate harmless duplicates; they are never stored.
// value of element e of a loaded row after dropout (identity without DROP)
template <bool DROP>
__device__ __forceinline__ float dropped(float v, uint32_t mk, int e, float scale) {
  if constexpr (DROP) return ((mk >> e) & 1u) ? __fmul_rn(v, scale) : 0.f;
  return v;
}

template <int VEC, int NCH, int U, bool IS_MAX, bool HAS_W, bool SQDEV, bool DROP>
__device__ __forceinline__ void accum_all(float (&acc)[NCH][VEC], int32_t (&aidx)[NCH][VEC],
                                          const float (&v)[U][NCH][VEC], const float (&w)[U], const int32_t (&c)[U],
                                          const bool (&on)[NCH], bool negate, const uint32_t (&mk)[U][NCH],
                                          float drop_scale) {
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        accum<IS_MAX, HAS_W, SQDEV>(acc[ch][e], aidx[ch][e], dropped<DROP>(v[u][ch][e], mk[u][ch], e, drop_scale), w[u],
                                    c[u], negate);
}

// Reduce CSR slots [k0, k1) of one row into acc (all lanes of the group call this together).
template <int VEC, int G, int NCH, bool IS_MAX, bool HAS_EW, bool HAS_SS, bool SPLIT, bool SQDEV, bool HOT, bool DROP>
__device__ __forceinline__ void reduce_range(const GRP& p, int64_t k0, int64_t k1, int gl,
                                             unsigned gmask, const bool (&on)[NCH],
                                             float (&acc)[NCH][VEC], int32_t (&aidx)[NCH][VEC]) {
  constexpr int UMAX = (8 / NCH) < 1 ? 1 : (8 / NCH);
  constexpr int U = (G < UMAX) ? G : UMAX;
  constexpr bool HAS_W = HAS_EW || HAS_SS;
  const bool negate = p.negate != 0;
  int64_t k = k0;
  int32_t myc = 0;
  int32_t mye = 0;
  float myw = 1.f;
  if (k + gl < k1) {
    if constexpr (DROP) mye = __ldg(p.edge_id + k + gl);
    myc = load_col<HOT>(p, k + gl);
    if constexpr (HAS_EW) myw = __ldg(p.edge_w + k + gl);
    if constexpr (HAS_SS) myw = __fmul_rn(myw, __ldg(p.src_scale + (myc & 0x7fffffff)));
  }
  while (k < k1) {
    const int64_t rem = k1 - k;
    const int cnt = rem < G ? (int)rem : G;
    const int64_t kn = k + G;
    int32_t nc = 0;
    int32_t ne = 0;
    float nw = 1.f;
    if (kn + gl < k1) {  // prefetch the next index batch while this one is consumed
      if constexpr (DROP) ne = __ldg(p.edge_id + kn + gl);
      nc = load_col<HOT>(p, kn + gl);
      if constexpr (HAS_EW) nw = __ldg(p.edge_w + kn + gl);
      if constexpr (HAS_SS) nw = __fmul_rn(nw, __ldg(p.src_scale + (nc & 0x7fffffff)));
    }
#pragma unroll 1
    for (int j = 0; j < cnt; j += U) {
      float v[U][NCH][VEC];
      float w[U];
      int32_t c[U];
      uint32_t mk[U][NCH];
      load_batch_full<VEC, G, NCH, U, HAS_W, SPLIT, HOT, DROP>(p, myc, myw, mye, j, gl, gmask, on, v, w, c, mk);
      if (j + U <= cnt) {  // group-uniform: a full batch needs no predication
        accum_all<VEC, NCH, U, IS_MAX, HAS_W, SQDEV, DROP>(acc, aidx, v, w, c, on, negate, mk, p.drop_scale);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if ((j + u) < cnt) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
              for (int e = 0; e < VEC; ++e)
                accum<IS_MAX, HAS_W, SQDEV>(acc[ch][e], aidx[ch][e], dropped<DROP>(v[u][ch][e], mk[u][ch], e, p.drop_scale), w[u], c[u], negate);
          }
        }
      }
    }
    myc = nc;
    mye = ne;
    myw = nw;
    k = kn;
  }
}

template <int VEC, int G, int NCH, bool IS_MAX, bool SPLIT, bool SQDEV>
__device__ __forceinline__ void epilogue(const GRP& p, int64_t slot, int64_t deg, int gl,
                                         const bool (&on)[NCH], float (&acc)[NCH][VEC],
                                         int32_t (&aidx)[NCH][VEC]) {
  const int64_t row_out = p.row_ids ? (int64_t)__ldg(p.row_ids + slot) : slot;
  float os = 1.f;
  if (p.out_scale) os = __ldg(p.out_scale + slot);
  if (SPLIT && row_out >= p.n_split_out) {
    // second output space: a local buffer, or (fused halo-gradient exchange) the owner's window over NVLink
    const int64_t r2 = row_out - p.n_split_out;
    float* drow = p.has_push ? push_row(p.tab, r2) : p.out2 + r2 * p.ldo2;
    const float den2 = fmaxf((float)deg, 1e-8f);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      if (!on[ch]) continue;
      float r[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        r[e] = acc[ch][e];
        if (p.mean) r[e] = __fdiv_rn(r[e], den2);
        if (p.out_scale) r[e] = __fmul_rn(r[e], os);
      }
      st_vec<VEC>(drow + (gl + ch * G) * VEC, r);
    }
    return;
  }
  // mean: sum / max(count, 1e-8) with a true IEEE division like the reference (aggregators.py:77-81);
  // it runs once per row and element, not per edge
  const float den = fmaxf((float)deg, 1e-8f);
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    if (!on[ch]) continue;
    const int f0 = (gl + ch * G) * VEC;
    float r[VEC];
    int32_t a[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      r[e] = acc[ch][e];
      a[e] = aidx[ch][e];
      if constexpr (IS_MAX) {
        if (p.negate) r[e] = -r[e];
        if (!p.raw_max && isinf(r[e])) { r[e] = 0.f; a[e] = -1; }  // empty segment, or a genuine +-inf (reference quirk)
        if (r[e] != r[e]) a[e] = -1;
      }
      if (p.mean) r[e] = __fdiv_rn(r[e], den);
      if constexpr (SQDEV) {
        // population std (aggregators.py:218-226): sqrt(max(sum_sq / max(count, 1e-8), 0)); count <= 1 -> 0
        r[e] = __fsqrt_rn(fmax_nan(__fdiv_rn(r[e], den), 0.f));   // NaN-propagating like torch.maximum
        if (deg <= 1) r[e] = 0.f;
      }
      if (p.out_scale) r[e] = __fmul_rn(r[e], os);
    }
    if (!SQDEV && p.addend) {
      float ad[VEC];
      ld_vec<VEC>(p.addend + row_out * p.ld_addend + f0, ad);
#pragma unroll
      for (int e = 0; e < VEC; ++e) r[e] = __fadd_rn(__fmul_rn(p.addend_scale, ad[e]), r[e]);
    }
    if (p.bias) {
      float b[VEC];
      ld_vec<VEC>(p.bias + f0, b);
#pragma unroll
      for (int e = 0; e < VEC; ++e) r[e] = __fadd_rn(r[e], b[e]);
    }
    if (p.act == KGB_ACT_RELU) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) r[e] = (r[e] <= 0.f) ? 0.f : r[e];  // keeps NaN like torch.relu
    }
    st_vec<VEC>(p.out + row_out * p.ldo + f0, r);
    if constexpr (IS_MAX) {
      if (p.arg) st_vec_i<VEC>(p.arg + row_out * p.ldo + f0, a);
    }
  }
}

// Walk the contiguous edge range of CSR rows [ra, rb) of the block starting at row r0 (group-uniform
// arguments).  my_hi holds rowptr[r0 + gl + 1] for lane gl of the group.
template <int VEC, int G, int NCH, bool IS_MAX, bool HAS_EW, bool HAS_SS, bool SPLIT, bool SQDEV, bool HOT, bool DROP>
__device__ __forceinline__ void walk_rows(const GRP& p, int64_t r0, int ra, int rb, int64_t k0, int64_t k1,
                                          int64_t my_hi, int gl, unsigned gmask, const bool (&on)[NCH]) {
  constexpr int UMAX = (8 / NCH) < 1 ? 1 : (8 / NCH);
  constexpr int U = (G < UMAX) ? G : UMAX;
  constexpr bool HAS_W = HAS_EW || HAS_SS;
  const bool negate = p.negate != 0;
  float acc[NCH][VEC];
  int32_t aidx[NCH][VEC];
  init_acc<VEC, G, NCH, IS_MAX>(acc, aidx);
  load_mu<VEC, G, NCH, SQDEV>(p, r0 + ra, gl, on, aidx);
  int cur = ra;
  int64_t cur_start = k0;
  int64_t cur_end = __shfl_sync(gmask, my_hi, cur, G);
  int64_t k = k0;
  int32_t myc = 0;
  int32_t mye = 0;
  float myw = 1.f;
  if (k + gl < k1) {
    if constexpr (DROP) mye = __ldg(p.edge_id + k + gl);
    myc = load_col<HOT>(p, k + gl);
    if constexpr (HAS_EW) myw = __ldg(p.edge_w + k + gl);
    if constexpr (HAS_SS) myw = __fmul_rn(myw, __ldg(p.src_scale + (myc & 0x7fffffff)));
  }
  while (k < k1) {
    const int64_t rem = k1 - k;
    const int cnt = rem < G ? (int)rem : G;
    const int64_t kn = k + G;
    int32_t nc = 0;
    int32_t ne = 0;
    float nw = 1.f;
    if (kn + gl < k1) {  // prefetch the next index batch while this one is consumed
      if constexpr (DROP) ne = __ldg(p.edge_id + kn + gl);
      nc = load_col<HOT>(p, kn + gl);
      if constexpr (HAS_EW) nw = __ldg(p.edge_w + kn + gl);
      if constexpr (HAS_SS) nw = __fmul_rn(nw, __ldg(p.src_scale + (nc & 0x7fffffff)));
    }
#pragma unroll 1
    for (int j = 0; j < cnt; j += U) {
      float v[U][NCH][VEC];
      float w[U];
      int32_t c[U];
      uint32_t mk[U][NCH];
      const int64_t kk0 = k + j;
      load_batch_full<VEC, G, NCH, U, HAS_W, SPLIT, HOT, DROP>(p, myc, myw, mye, j, gl, gmask, on, v, w, c, mk);
      if (j + U <= cnt && kk0 + U <= cur_end) {
        // fast path (group-uniform): a full batch that lies inside the current row
        accum_all<VEC, NCH, U, IS_MAX, HAS_W, SQDEV, DROP>(acc, aidx, v, w, c, on, negate, mk, p.drop_scale);
        continue;
      }
      // distribute the (up to U) loaded edges over the rows they belong to
      const int valid = (cnt - j) < U ? (cnt - j) : U;
      int done = 0;
      while (true) {
        const int64_t left = cur_end - (kk0 + done);  // edges of the current row still ahead
        const int lim = (left < (int64_t)(valid - done)) ? done + (int)left : valid;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (u >= done && u < lim) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
              for (int e = 0; e < VEC; ++e)
                accum<IS_MAX, HAS_W, SQDEV>(acc[ch][e], aidx[ch][e], dropped<DROP>(v[u][ch][e], mk[u][ch], e, p.drop_scale), w[u], c[u], negate);
          }
        }
        done = lim;
        if (done >= valid) break;
        // the current row is complete: flush it and move to the next one
        epilogue<VEC, G, NCH, IS_MAX, SPLIT, SQDEV>(p, r0 + cur, cur_end - cur_start, gl, on, acc, aidx);
        init_acc<VEC, G, NCH, IS_MAX>(acc, aidx);
        ++cur;
        load_mu<VEC, G, NCH, SQDEV>(p, r0 + cur, gl, on, aidx);
        cur_start = cur_end;
        cur_end = __shfl_sync(gmask, my_hi, cur, G);
      }
    }
    myc = nc;
    mye = ne;
    myw = nw;
    k = kn;
  }
  // rows that end exactly at k1 (the last one with edges, then any empty rows)
  while (cur < rb) {
    epilogue<VEC, G, NCH, IS_MAX, SPLIT, SQDEV>(p, r0 + cur, cur_end - cur_start, gl, on, acc, aidx);
    init_acc<VEC, G, NCH, IS_MAX>(acc, aidx);
    ++cur;
    cur_start = cur_end;
    if (cur < rb) cur_end = __shfl_sync(gmask, my_hi, cur, G);   // the remaining rows are empty: no mean needed
  }
}


// Rows handed out per queue fetch (one atomic per warp per UNIT_ROWS rows).
#ifndef KGB_GR_UNIT_ROWS
#define KGB_GR_UNIT_ROWS 32
#endif
constexpr int UNIT_ROWS = KGB_GR_UNIT_ROWS;

template <int VEC, int G, int NCH, bool IS_MAX, bool HAS_EW, bool HAS_SS, bool SPLIT, bool SQDEV, bool HOT, bool DROP>
__device__ __forceinline__ void do_chunk(const GRP& p, int64_t t, int gl, unsigned gmask, const bool (&on)[NCH]) {
  // one chunk of a hub row -> raw partial
  float acc[NCH][VEC];
  int32_t aidx[NCH][VEC];
  init_acc<VEC, G, NCH, IS_MAX>(acc, aidx);
  const int h = __ldg(p.chunk_hub + t);
  const int64_t row = __ldg(p.hub_row + h);
  load_mu<VEC, G, NCH, SQDEV>(p, row, gl, on, aidx);
  const int64_t ci = t - __ldg(p.hub_chunk_base + h);
  const int64_t rs = __ldg(p.rowptr + row), re = __ldg(p.rowptr + row + 1);
  const int64_t k0 = rs + ci * p.hub_chunk;
  const int64_t k1 = (k0 + p.hub_chunk < re) ? k0 + p.hub_chunk : re;
  reduce_range<VEC, G, NCH, IS_MAX, HAS_EW, HAS_SS, SPLIT, SQDEV, HOT, DROP>(p, k0, k1, gl, gmask, on, acc, aidx);
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    if (!on[ch]) continue;
    const int f0 = (gl + ch * G) * VEC;
    st_vec<VEC>(p.partial + t * (int64_t)p.F + f0, acc[ch]);
    if constexpr (IS_MAX) st_vec_i<VEC>(p.partial_arg + t * (int64_t)p.F + f0, aidx[ch]);
  }
}

template <int VEC, int G, int NCH, bool IS_MAX, bool HAS_EW, bool HAS_SS, bool SPLIT, bool SQDEV, bool HOT, bool DROP>
__device__ __forceinline__ void do_row_block(const GRP& p, int64_t r0, int gl, int gw, unsigned gmask,
                                             const bool (&on)[NCH]) {
  // a block of G consecutive rows: lane gl holds rowptr[r0+gl], rowptr[r0+gl+1]
  const int64_t left = p.n_rows - r0;
  const int nr = left < G ? (int)left : G;
  int64_t my_lo = 0, my_hi = 0;
  if (gl < nr) {
    my_lo = __ldg(p.rowptr + r0 + gl);
    my_hi = __ldg(p.rowptr + r0 + gl + 1);
  }
  const bool is_hub = (p.n_hubs > 0) && (gl < nr) && ((my_hi - my_lo) > p.hub_threshold);
  const unsigned hubmask = (__ballot_sync(gmask, is_hub) & gmask) >> (gw * G);  // group-relative bits
  int cur = 0;
  while (cur < nr) {
    const unsigned m = hubmask >> cur;
    const int seg_end = m ? cur + __ffs(m) - 1 : nr;  // first hub row at or after cur
    if (seg_end > cur) {
      const int64_t k0 = __shfl_sync(gmask, my_lo, cur, G);
      const int64_t k1 = __shfl_sync(gmask, my_hi, seg_end - 1, G);
      walk_rows<VEC, G, NCH, IS_MAX, HAS_EW, HAS_SS, SPLIT, SQDEV, HOT, DROP>(p, r0, cur, seg_end, k0, k1, my_hi, gl, gmask, on);
    }
    cur = seg_end + 1;  // the hub row (if any) is written by hub_finish_kernel
  }
}

// resident CTAs per SM the register allocation is tuned for (8 float4 loads in flight per lane)
#ifndef KGB_GR_MINB_WIDE
#define KGB_GR_MINB_WIDE 3  // VEC == 4 and G == 32 (rows of >= 17 vectors): the HBM-bound shapes.  4 CTAs/SM were measured on
                            // C4: F=48 (G=16) 2.02 / 2.58 -> 1.97 / 2.41 ms fwd / bwd (latency-bound, so G=16 is "narrow"),
                            // F=100 2.73 / 3.42 -> 2.68 / 3.62, F=256 6.02 / 6.59 -> 6.51 / 7.58 (spills)
#endif
#ifndef KGB_GR_MINB_NARROW
#define KGB_GR_MINB_NARROW 4
#endif
#ifndef KGB_GR_MINB_MAX
#define KGB_GR_MINB_MAX 3  // max/min carry value + argmax per element (twice the accumulator registers)
#endif

// Tasks: first the hub chunks (the long ones), then units of UNIT_ROWS consecutive rows.  They are
// handed out by a global queue (one atomicAdd per warp and unit) because static striding correlates
// with the id structure of power-law graphs (RMAT: the degree depends on the low id bits), which left
// 27 % of the SM-cycles idle in the first version.  Which warp computes a row never changes the
// result, so the output stays deterministic.
template <int VEC, int G, int NCH, bool IS_MAX, bool HAS_EW, bool HAS_SS, bool SPLIT, bool SQDEV, bool HOT, bool DROP>
__global__ void __launch_bounds__(256, ((VEC == 4 && G >= 32) ? (IS_MAX ? KGB_GR_MINB_MAX : KGB_GR_MINB_WIDE)
                                                              : KGB_GR_MINB_NARROW))
gather_reduce_kernel(const GRP p) {
  constexpr int GPW = 32 / G;
  constexpr int BPU = UNIT_ROWS / G;  // row blocks per unit
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int gw = lane / G;
  const unsigned gmask = group_mask(lane, G);
  const int nv = p.F / VEC;
  bool on[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) on[ch] = (gl + ch * G) < nv;

  const int64_t chunk_units = ((int64_t)p.n_chunks + GPW - 1) / GPW;
  const int64_t row_units = (p.n_rows + UNIT_ROWS - 1) / UNIT_ROWS;
  const int64_t n_units = chunk_units + row_units;
  const int64_t wpb = blockDim.x >> 5;
  int64_t u = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5);  // static fallback when no queue is given
  while (true) {
    if (p.work) {
      __syncwarp();
      int32_t t = 0;
      if (lane == 0) t = atomicAdd(p.work, 1);
      u = __shfl_sync(0xffffffffu, t, 0);
    }
    if (u >= n_units) break;
    if (u < chunk_units) {
      const int64_t t = u * GPW + gw;
      if (t < p.n_chunks) do_chunk<VEC, G, NCH, IS_MAX, HAS_EW, HAS_SS, SPLIT, SQDEV, HOT, DROP>(p, t, gl, gmask, on);
    } else {
      int64_t ui = u - chunk_units;
      if (p.unit_order) ui = __ldg(p.unit_order + ui);
      const int64_t base = ui * UNIT_ROWS;
      for (int b = gw; b < BPU; b += GPW) {
        const int64_t r0 = base + (int64_t)b * G;
        if (r0 < p.n_rows) do_row_block<VEC, G, NCH, IS_MAX, HAS_EW, HAS_SS, SPLIT, SQDEV, HOT, DROP>(p, r0, gl, gw, gmask, on);
      }
    }
    if (!p.work) u += (int64_t)gridDim.x * wpb;
  }
  if (p.work) {  // the last CTA to finish re-arms the queue for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int done = atomicAdd(p.work + 1, 1);
      if (done == (int)gridDim.x - 1) {
        p.work[0] = 0;
        p.work[1] = 0;
        __threadfence();
      }
    }
  }
}

// Merge the chunk partials of every hub row in chunk order, then run the epilogue.
template <int VEC, int G, int NCH, bool IS_MAX, bool SPLIT, bool SQDEV>
__global__ void __launch_bounds__(256) hub_finish_kernel(const GRP p) {
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int gw = lane / G;
  const int nv = p.F / VEC;
  bool on[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) on[ch] = (gl + ch * G) < nv;
  const int64_t gpb = (int64_t)(blockDim.x >> 5) * GPW;
  for (int64_t h = (int64_t)blockIdx.x * gpb + (int64_t)(threadIdx.x >> 5) * GPW + gw; h < p.n_hubs;
       h += (int64_t)gridDim.x * gpb) {
    float acc[NCH][VEC];
    int32_t aidx[NCH][VEC];
    init_acc<VEC, G, NCH, IS_MAX>(acc, aidx);
    const int64_t row = __ldg(p.hub_row + h);
    const int64_t base = __ldg(p.hub_chunk_base + h);
    const int nch = __ldg(p.hub_nchunks + h);
    // chunk partials are merged in chunk order (fixed summation order); HU of them are loaded ahead of the adds
    // so the merge of a 500-chunk hub is not one dependent round trip per chunk
    constexpr int HU = (NCH >= 4 || IS_MAX) ? 2 : 4;
    for (int c0 = 0; c0 < nch; c0 += HU) {
      float v[HU][NCH][VEC];
      int32_t pa[HU][NCH][VEC];
#pragma unroll
      for (int u = 0; u < HU; ++u) {
        const int64_t c = (c0 + u < nch) ? (c0 + u) : (nch - 1);   // clamped (re-read, not re-added)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          if (!on[ch]) continue;
          const int f0 = (gl + ch * G) * VEC;
          ld_vec<VEC>(p.partial + (base + c) * (int64_t)p.F + f0, v[u][ch]);
          if constexpr (IS_MAX) ld_vec_i<VEC>(p.partial_arg + (base + c) * (int64_t)p.F + f0, pa[u][ch]);
        }
      }
#pragma unroll
      for (int u = 0; u < HU; ++u) {
        if (c0 + u >= nch) break;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          if (!on[ch]) continue;
          if constexpr (IS_MAX) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              float& m = acc[ch][e];
              const float ve = v[u][ch][e];
              const int32_t pe = pa[u][ch][e];
              if (ve != ve) m = ve;
              else if (ve > m) { m = ve; aidx[ch][e] = pe; }
              else if (ve == m && pe != -1 && pe != aidx[ch][e]) aidx[ch][e] = -2;
            }
          } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[ch][e] = __fadd_rn(acc[ch][e], v[u][ch][e]);
          }
        }
      }
    }
    const int64_t rs = __ldg(p.rowptr + row), re = __ldg(p.rowptr + row + 1);
    epilogue<VEC, G, NCH, IS_MAX, SPLIT, SQDEV>(p, row, re - rs, gl, on, acc, aidx);
  }
}

// ---- backward of max/min -------------------------------------------------------------------
// torch.scatter_reduce(amax) backward: g[r,f] goes to the source attaining the extremum; when several
// DIFFERENT sources tie (arg == -2) it is split evenly over every tied edge.  Unique entries are an
// N*F scatter (no edge traffic).  Tied entries re-walk the row with batched index loads; tied entries
// of hub rows are handled chunk-parallel by three small follow-up kernels (count, total, scatter).
struct MaxBwdP {
  const float* g; int64_t ldg; const int32_t* arg; const float* out; int64_t ldo;
  const float* x; int64_t ldx; const int64_t* rowptr; const int32_t* col; const int32_t* row_ids;
  int64_t n_rows; int F; float* gx; int64_t ldgx;
  const int32_t* hub_row; const int32_t* hub_chunk_base; const int32_t* hub_nchunks; const int32_t* chunk_hub;
  int n_hubs; int n_chunks; int hub_threshold; int hub_chunk;
  float* cnt_partial;  // [n_chunks, F]
  float* cnt_total;    // [n_hubs, F]
  int32_t* hub_tie;    // [n_hubs]
  // deterministic scatter: contributions are added as 64-bit fixed-point integers (integer addition is associative,
  // so the order in which the atomics land cannot change the result); gmax_bits = bit pattern of max |g| over the
  // finite entries, written by max_bwd_absmax_kernel before the scatter kernels run
  long long* acc; int64_t ldacc; uint32_t* gmax_bits; int cnt_bits;  // gmax_bits[1]: a non-finite gradient was seen
};

// exponent s of the fixed-point grid 2^-s: max|g| < 2^e and at most 2^cnt_bits contributions of magnitude <= max|g|
// reach one slot, so |sum * 2^s| < 2^62
__device__ __forceinline__ int max_bwd_shift(const MaxBwdP& p) {
  const float gm = __uint_as_float(__ldg(p.gmax_bits));
  if (!(gm > 0.f)) return 0;
  int e;
  frexpf(gm, &e);  // gm = m * 2^e, 0.5 <= m < 1
  return 62 - e - p.cnt_bits;
}

// gx[row, f] += v, deterministically.  Non-finite contributions (NaN / +-inf gradients) go to the fp32 buffer with a
// float atomic: any order of a multiset of NaN / inf values gives the same result.
__device__ __forceinline__ void max_bwd_add(const MaxBwdP& p, int shift, int64_t row, int f, float v) {
  if (v == 0.f) return;
  if (fabsf(v) <= 3.4028234e38f) {
    const long long q = __double2ll_rn(ldexp((double)v, shift));
    atomicAdd(reinterpret_cast<unsigned long long*>(p.acc + row * p.ldacc + f), (unsigned long long)q);
  } else {
    atomicAdd(p.gx + row * p.ldgx + f, v);
    p.gmax_bits[1] = 1u;
  }
}

// Walk CSR slots [k0, k1): phase 0 counts, per element, the edges whose source value equals o; phase 1
// adds gv / cnt to those sources.  Only elements with a == -2 take part.
template <int VEC, int G, int NCH, int PHASE>
__device__ __forceinline__ void tie_walk(const MaxBwdP& p, int64_t k0, int64_t k1, int gl, unsigned gmask,
                                         const bool (&on)[NCH], const int32_t (&a)[NCH][VEC],
                                         const float (&o)[NCH][VEC], const float (&gv)[NCH][VEC],
                                         float (&cnt)[NCH][VEC], int shift = 0) {
  constexpr int U = (G < 4) ? G : 4;
  int loff[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) loff[ch] = on[ch] ? (gl + ch * G) * VEC : 0;
  for (int64_t k = k0; k < k1; k += G) {
    const int64_t rem = k1 - k;
    const int cnt_e = rem < G ? (int)rem : G;
    const int32_t myc = (k + gl < k1) ? __ldg(p.col + k + gl) : 0;
#pragma unroll 1
    for (int j = 0; j < cnt_e; j += U) {
      float v[U][NCH][VEC];
      int32_t c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        c[u] = __shfl_sync(gmask, myc, j + u, G);
        const float* rp = p.x + (int64_t)c[u] * p.ldx;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) ld_vec<VEC>(rp + loff[ch], v[u][ch]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if ((j + u) >= cnt_e) continue;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          if (!on[ch]) continue;
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            // special entries: ties between different sources (arg == -2) and non-finite gradients, which the
            // reference multiplies with a 0/1 mask over ALL edges of the row (0 * inf = NaN reaches every source)
            const bool nonfin = !(fabsf(gv[ch][e]) <= 3.4028234e38f);
            const bool special = a[ch][e] == -2 || (a[ch][e] >= 0 && nonfin);
            if (!special) continue;
            if (v[u][ch][e] == o[ch][e]) {
              if (PHASE == 0) cnt[ch][e] += 1.f;
              else max_bwd_add(p, shift, (int64_t)c[u], (gl + ch * G) * VEC + e, __fdiv_rn(gv[ch][e], cnt[ch][e]));
            } else if (PHASE == 1 && nonfin) {
              max_bwd_add(p, shift, (int64_t)c[u], (gl + ch * G) * VEC + e, __fmul_rn(gv[ch][e], 0.f));
            }
          }
        }
      }
    }
  }
}

template <int VEC, int G, int NCH>
__device__ __forceinline__ bool max_bwd_load_row(const MaxBwdP& p, int shift, int64_t r, int gl, unsigned gmask,
                                                 const bool (&on)[NCH], float (&gv)[NCH][VEC], int32_t (&a)[NCH][VEC]) {
  bool tie = false;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) { gv[ch][e] = 0.f; a[ch][e] = -1; }
    if (!on[ch]) continue;
    const int f0 = (gl + ch * G) * VEC;
    ld_vec<VEC>(p.g + r * p.ldg + f0, gv[ch]);
    ld_vec_i<VEC>(p.arg + r * p.ldo + f0, a[ch]);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const bool nonfin = !(fabsf(gv[ch][e]) <= 3.4028234e38f);
      if (a[ch][e] >= 0 && !nonfin) max_bwd_add(p, shift, (int64_t)a[ch][e], f0 + e, gv[ch][e]);
      tie |= (a[ch][e] == -2) || (a[ch][e] >= 0 && nonfin);   // both re-walk the row
    }
  }
  return (__ballot_sync(gmask, tie) & gmask) != 0;
}

// pass 1: every row scatters its unique entries; non-hub rows with ties re-walk themselves; hub rows
// (mode 1, one group per hub) only raise hub_tie[h]
template <int VEC, int G, int NCH, int MODE>
__global__ void __launch_bounds__(256) gather_max_bwd_kernel(const MaxBwdP p) {
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int gw = lane / G;
  const unsigned gmask = group_mask(lane, G);
  const int nv = p.F / VEC;
  bool on[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) on[ch] = (gl + ch * G) < nv;
  const int64_t gpb = (int64_t)(blockDim.x >> 5) * GPW;
  const int64_t n_items = MODE == 0 ? p.n_rows : (int64_t)p.n_hubs;
  const int shift = max_bwd_shift(p);
  for (int64_t it = (int64_t)blockIdx.x * gpb + (int64_t)(threadIdx.x >> 5) * GPW + gw; it < n_items;
       it += (int64_t)gridDim.x * gpb) {
    const int64_t s = MODE == 0 ? it : (int64_t)__ldg(p.hub_row + it);
    const int64_t rs = __ldg(p.rowptr + s), re = __ldg(p.rowptr + s + 1);
    if (MODE == 0 && p.n_hubs > 0 && (re - rs) > p.hub_threshold) continue;  // hub rows: mode-1 launch
    const int64_t r = p.row_ids ? (int64_t)__ldg(p.row_ids + s) : s;
    float gv[NCH][VEC];
    int32_t a[NCH][VEC];
    const bool any_tie = max_bwd_load_row<VEC, G, NCH>(p, shift, r, gl, gmask, on, gv, a);
    if (!any_tie) continue;
    if (MODE == 1) {
      if (gl == 0) p.hub_tie[it] = 1;
      continue;
    }
    float o[NCH][VEC], cnt[NCH][VEC];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) { o[ch][e] = 0.f; cnt[ch][e] = 0.f; }
      if (on[ch]) ld_vec<VEC>(p.out + r * p.ldo + (gl + ch * G) * VEC, o[ch]);
    }
    tie_walk<VEC, G, NCH, 0>(p, rs, re, gl, gmask, on, a, o, gv, cnt);
    tie_walk<VEC, G, NCH, 1>(p, rs, re, gl, gmask, on, a, o, gv, cnt, shift);
  }
}

// pass 2 (PHASE 0): per chunk of a hub row that has ties, count the tied edges -> cnt_partial[t,:]
// pass 4 (PHASE 1): per chunk, scatter g / cnt_total to the tied edges
template <int VEC, int G, int NCH, int PHASE>
__global__ void __launch_bounds__(256) max_bwd_hub_chunk_kernel(const MaxBwdP p) {
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int gw = lane / G;
  const unsigned gmask = group_mask(lane, G);
  const int nv = p.F / VEC;
  bool on[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) on[ch] = (gl + ch * G) < nv;
  const int64_t gpb = (int64_t)(blockDim.x >> 5) * GPW;
  const int shift = (PHASE == 1) ? max_bwd_shift(p) : 0;
  for (int64_t t = (int64_t)blockIdx.x * gpb + (int64_t)(threadIdx.x >> 5) * GPW + gw; t < p.n_chunks;
       t += (int64_t)gridDim.x * gpb) {
    const int h = __ldg(p.chunk_hub + t);
    if (p.hub_tie[h] == 0) continue;
    const int64_t s = __ldg(p.hub_row + h);
    const int64_t r = p.row_ids ? (int64_t)__ldg(p.row_ids + s) : s;
    const int64_t ci = t - __ldg(p.hub_chunk_base + h);
    const int64_t rs = __ldg(p.rowptr + s), re = __ldg(p.rowptr + s + 1);
    const int64_t k0 = rs + ci * p.hub_chunk;
    const int64_t k1 = (k0 + p.hub_chunk < re) ? k0 + p.hub_chunk : re;
    float gv[NCH][VEC], o[NCH][VEC], cnt[NCH][VEC];
    int32_t a[NCH][VEC];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) { gv[ch][e] = 0.f; o[ch][e] = 0.f; cnt[ch][e] = 0.f; a[ch][e] = -1; }
      if (!on[ch]) continue;
      const int f0 = (gl + ch * G) * VEC;
      ld_vec<VEC>(p.g + r * p.ldg + f0, gv[ch]);
      ld_vec_i<VEC>(p.arg + r * p.ldo + f0, a[ch]);
      ld_vec<VEC>(p.out + r * p.ldo + f0, o[ch]);
      if (PHASE == 1) ld_vec<VEC>(p.cnt_total + (int64_t)h * p.F + f0, cnt[ch]);
    }
    tie_walk<VEC, G, NCH, PHASE>(p, k0, k1, gl, gmask, on, a, o, gv, cnt, shift);
    if (PHASE == 0) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
        if (on[ch]) st_vec<VEC>(p.cnt_partial + t * (int64_t)p.F + (gl + ch * G) * VEC, cnt[ch]);
    }
  }
}

// pass 3: cnt_total[h,:] = sum over the hub's chunks of cnt_partial (only hubs with ties)
__global__ void __launch_bounds__(256) max_bwd_hub_total_kernel(const MaxBwdP p) {
  const int64_t total = (int64_t)p.n_hubs * p.F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(i / p.F), f = (int)(i % p.F);
    if (p.hub_tie[h] == 0) continue;
    const int64_t base = __ldg(p.hub_chunk_base + h);
    const int nch = __ldg(p.hub_nchunks + h);
    float s = 0.f;
    for (int c = 0; c < nch; ++c) s += p.cnt_partial[(base + c) * p.F + f];
    p.cnt_total[i] = s;
  }
}

// max |g| over the finite entries of g [rows, F] (bit patterns of non-negative floats order like unsigned integers).
// One warp per row (grid-stride), 128-bit loads when the rows allow it; no per-element index arithmetic.
__device__ __forceinline__ uint32_t absbits_finite(float v, uint32_t m) {
  const uint32_t b = __float_as_uint(v) & 0x7fffffffu;
  return (b < 0x7f800000u && b > m) ? b : m;
}

__global__ void __launch_bounds__(256) max_bwd_absmax_kernel(const float* __restrict__ g, int64_t ldg, int64_t rows,
                                                             int F, int vec4, uint32_t* __restrict__ out_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  uint32_t m = 0;
  for (int64_t r = warp; r < rows; r += n_warps) {
    const float* row = g + r * ldg;
    if (vec4) {
      for (int f = lane * 4; f < F; f += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + f));
        m = absbits_finite(v.x, m); m = absbits_finite(v.y, m); m = absbits_finite(v.z, m); m = absbits_finite(v.w, m);
      }
    } else {
      for (int f = lane; f < F; f += 32) m = absbits_finite(__ldg(row + f), m);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if (lane == 0 && m) atomicMax(out_bits, m);
}

// gx += acc * 2^-shift (the fixed-point sums back to fp32, one correctly rounded conversion per element)
__global__ void __launch_bounds__(256) max_bwd_convert_kernel(const MaxBwdP p, int64_t n_src) {
  const int shift = max_bwd_shift(p);
  const bool nonfinite = p.gmax_bits[1] != 0u;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool pair = (p.F % 2 == 0) && (p.ldacc % 2 == 0) && (p.ldgx % 2 == 0) &&
                    ((reinterpret_cast<uintptr_t>(p.acc) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(p.gx) & 7u) == 0);
  for (int64_t r = warp; r < n_src; r += n_warps) {   // one warp per row: no per-element index arithmetic
    const long long* arow = p.acc + r * p.ldacc;
    float* grow = p.gx + r * p.ldgx;
    if (pair) {
      for (int f = lane * 2; f < p.F; f += 64) {
        const longlong2 q = *reinterpret_cast<const longlong2*>(arow + f);
        if ((q.x | q.y) != 0) {  // gx is zero on entry except where non-finite gradients landed (flagged, rare)
          float2 v = make_float2((float)ldexp((double)q.x, -shift), (float)ldexp((double)q.y, -shift));
          if (nonfinite) {
            const float2 o = *reinterpret_cast<const float2*>(grow + f);
            v.x = __fadd_rn(o.x, v.x); v.y = __fadd_rn(o.y, v.y);
          }
          *reinterpret_cast<float2*>(grow + f) = v;
        }
      }
    } else {
      for (int f = lane; f < p.F; f += 32) {
        const long long q = arow[f];
        if (q != 0) {
          const float v = (float)ldexp((double)q, -shift);
          grow[f] = nonfinite ? __fadd_rn(grow[f], v) : v;
        }
      }
    }
  }
}

template <int VEC, int G, int NCH>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ idx,
                   int64_t n_out, int F, float scale, float* __restrict__ out, int64_t ldo) {
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int gw = lane / G;
  const int nv = F / VEC;
  const int64_t gpb = (int64_t)(blockDim.x >> 5) * GPW;
  for (int64_t r = (int64_t)blockIdx.x * gpb + (int64_t)(threadIdx.x >> 5) * GPW + gw; r < n_out;
       r += (int64_t)gridDim.x * gpb) {
    const int64_t s = idx ? (int64_t)__ldg(idx + r) : r;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      if ((gl + ch * G) >= nv) continue;
      const int f0 = (gl + ch * G) * VEC;
      float v[VEC];
      ld_vec<VEC>(src + s * lds + f0, v);
      if (scale != 1.f) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = __fmul_rn(v[e], scale);
      }
      st_vec<VEC>(out + r * ldo + f0, v);
    }
  }
}

// K7: pack + push.  Slot s (grouped by destination peer) reads one feature row of this rank and stores it into the
// destination rank's window over NVLink (peer memory mapped with CUDA IPC).
template <int VEC, int G, int NCH>
__global__ void __launch_bounds__(256)
halo_push_kernel(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ idx, int F, int64_t ldd,
                 const PushTab tab) {
  constexpr int GPW = 32 / G;
  constexpr int U = NCH >= 4 ? 2 : 4;   // rows in flight per lane group: local HBM reads ahead of the NVLink stores
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int gw = lane / G;
  const int nv = F / VEC;
  const int64_t n_slots = tab.slot_begin[tab.n_peers];
  const int64_t n_walk = tab.n_chunks * PUSH_CHUNK;
  const int64_t gpb = (int64_t)(blockDim.x >> 5) * GPW;
  const int64_t stride = (int64_t)gridDim.x * gpb;
  for (int64_t s0 = (int64_t)blockIdx.x * gpb + (int64_t)(threadIdx.x >> 5) * GPW + gw; s0 < n_walk;
       s0 += (int64_t)U * stride) {
    int64_t r[U];
    float v[U][NCH][VEC];
    int64_t sl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t sw = s0 + (int64_t)u * stride;          // position in the walk
      const int64_t k = sw / PUSH_CHUNK;
      const int64_t kp = (int64_t)(((unsigned long long)k * (unsigned long long)tab.perm_mul +
                                    (unsigned long long)tab.perm_add) % (unsigned long long)tab.n_chunks);
      const int64_t s = kp * PUSH_CHUNK + (sw - k * PUSH_CHUNK);
      sl[u] = s;
      r[u] = (sw < n_walk && s < n_slots) ? (idx ? (int64_t)__ldg(idx + s) : s) : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r[u] < 0) continue;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
        if ((gl + ch * G) < nv) ld_vec<VEC>(src + r[u] * lds + (gl + ch * G) * VEC, v[u][ch]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r[u] < 0) continue;
      float* drow = push_row(tab, sl[u]);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
        if ((gl + ch * G) < nv) st_vec<VEC>(drow + (gl + ch * G) * VEC, v[u][ch]);
    }
  }
}

// gm = g where y > 0 else 0 (backward of a fused ReLU epilogue), one pass, 128-bit accesses when aligned
__global__ void relu_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ y, int64_t ldy,
                                int64_t rows, int F, float* __restrict__ out) {
  const int64_t total = rows * (int64_t)F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F;
    const int f = (int)(i - r * F);
    out[i] = y[r * ldy + f] > 0.f ? g[r * ldg + f] : 0.f;
  }
}

__global__ void relu_bwd_vec4_kernel(const float4* __restrict__ g, const float4* __restrict__ y, int64_t n4,
                                     float4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = g[i], b = y[i];
    out[i] = make_float4(b.x > 0.f ? a.x : 0.f, b.y > 0.f ? a.y : 0.f, b.z > 0.f ? a.z : 0.f, b.w > 0.f ? a.w : 0.f);
  }
}

__global__ void permute_f32_kernel(const float* __restrict__ in, const int32_t* __restrict__ perm,
                                   int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __ldg(in + __ldg(perm + i));
}

__global__ void reduce_parts_kernel(const float* __restrict__ part, int n_parts, int F, float* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  float s = 0.f;
  for (int p = 0; p < n_parts; ++p) s = __fadd_rn(s, part[(int64_t)p * F + f]);
  out[f] = s;
}

// keep an element when its 32 random bits are >= p * 2^32
static inline uint32_t dropout_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (t < 1.0 ? 1u : (uint32_t)t);
}

// the dropout mask the kernels apply, written out for tests: out[e, f] = 1 / (1 - p) if element f of edge e survives
__global__ void dropout_mask_kernel(const int32_t* __restrict__ edge_id, int64_t n_edges, int F, int per_head,
                                    uint32_t thr, float scale, uint32_t seed_lo, uint32_t seed_hi,
                                    float* __restrict__ out) {
  const int64_t total = n_edges * (int64_t)F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / F;
    const uint32_t f = (uint32_t)(i - e * F);
    const uint32_t eid = edge_id ? (uint32_t)__ldg(edge_id + e) : (uint32_t)e;
    const uint32_t bits = dropout_keep4(eid, f >> 2, per_head ? 1u : 0u, seed_lo, seed_hi, thr);
    out[i] = ((bits >> (f & 3u)) & 1u) ? scale : 0.f;
  }
}

// ---- host-side dispatch ----------------------------------------------------------------------
struct Shape {
  int vec, g, nch;
};

// feature slab handled by one launch: at most 32 lanes * 4 chunks * VEC floats
static Shape pick_shape(int F, bool can_vec4) {
  Shape s;
  s.vec = (can_vec4 && F % 4 == 0) ? 4 : 1;
  const int nv = F / s.vec;
  // (Rows of 9..32 vectors on half as many lanes with two vectors per lane - half the index / address work per loaded
  //  byte, more groups per warp - were measured slower on C4: F = 48 mean 2.04 -> 2.63 ms, F = 100 2.75 -> 3.78 ms;
  //  the extra groups per warp diverge on the power-law row lengths.)
  if (nv <= 32) {
    s.g = pow2_ceil(nv);
    s.nch = 1;
  } else {
    s.g = 32;
    const int n = (int)ceil_div(nv, 32);
    s.nch = n <= 2 ? 2 : 4;
  }
  return s;
}

static int grid_for(int device, int64_t tasks, int g) {
  const int64_t gpb = 8 * (32 / g);
  int64_t need = ceil_div(tasks, gpb);
  const int64_t cap = (int64_t)sm_count(device) * 8;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

template <int VEC, int G, int NCH>
static int launch_gr(int device, const GRP& p, bool is_max, cudaStream_t st) {
  // one warp per queue unit at a time; enough CTAs to fill every SM at the tuned residency
  const int64_t units = ceil_div((int64_t)p.n_chunks, 32 / G) + ceil_div(p.n_rows, (int64_t)UNIT_ROWS);
  int64_t need = ceil_div(units, 8);
  const int64_t cap = (int64_t)sm_count(device) * 4;
  const int grid = (int)(need > cap ? cap : (need < 1 ? 1 : need));
  const bool ew = p.edge_w != nullptr, ss = p.src_scale != nullptr;
  // the split-source / split-output variant is a separate instantiation: the common path pays nothing for it
  const bool split = p.x2 != nullptr || p.n_split_out != INT64_MAX;
#define KGB_GR_LAUNCH(MAXF, EW, SS)                                                                            \
  do {                                                                                                         \
    if (split) gather_reduce_kernel<VEC, G, NCH, MAXF, EW, SS, !MAXF, false, false, false><<<grid, 256, 0, st>>>(p);  \
    else if (drop && !EW && !MAXF) gather_reduce_kernel<VEC, G, NCH, false, false, SS, false, false, false, true><<<grid, 256, 0, st>>>(p); \
    else if (hot && !EW) gather_reduce_kernel<VEC, G, NCH, MAXF, false, SS, false, false, true, false><<<grid, 256, 0, st>>>(p); \
    else gather_reduce_kernel<VEC, G, NCH, MAXF, EW, SS, false, false, false, false><<<grid, 256, 0, st>>>(p); \
  } while (0)
  const bool drop = p.drop_thr != 0u;
  if (drop && (is_max || ew || split || p.sqdev)) {
    set_error("in-kernel dropout is available for sum / mean without per-edge weights or split operands");
    return KGB_ERR_INVALID;
  }
  const bool hot = p.col_hot != nullptr && !split;
  if (p.sqdev) {
    if (is_max || ew || ss || split) { set_error("the squared-deviation pass takes no weights / split operands"); return KGB_ERR_INVALID; }
    gather_reduce_kernel<VEC, G, NCH, false, false, false, false, true, false, false><<<grid, 256, 0, st>>>(p);
  } else if (is_max) {
    if (ew || ss) { set_error("max/min do not take edge weights"); return KGB_ERR_INVALID; }
    if (split) { set_error("max/min do not take split operands"); return KGB_ERR_INVALID; }
    KGB_GR_LAUNCH(true, false, false);
  } else if (ew && ss) {
    KGB_GR_LAUNCH(false, true, true);
  } else if (ew) {
    KGB_GR_LAUNCH(false, true, false);
  } else if (ss) {
    KGB_GR_LAUNCH(false, false, true);
  } else {
    KGB_GR_LAUNCH(false, false, false);
  }
#undef KGB_GR_LAUNCH
  KGB_CHECK_LAUNCH();
  if (p.n_hubs > 0) {
    const int hgrid = grid_for(device, p.n_hubs, G);
    if (p.sqdev) hub_finish_kernel<VEC, G, NCH, false, false, true><<<hgrid, 256, 0, st>>>(p);
    else if (is_max) hub_finish_kernel<VEC, G, NCH, true, false, false><<<hgrid, 256, 0, st>>>(p);
    else if (split) hub_finish_kernel<VEC, G, NCH, false, true, false><<<hgrid, 256, 0, st>>>(p);
    else hub_finish_kernel<VEC, G, NCH, false, false, false><<<hgrid, 256, 0, st>>>(p);
    KGB_CHECK_LAUNCH();
  }
  return KGB_OK;
}

#define KGB_DISPATCH_SHAPE(S, CALL)                                        \
  do {                                                                     \
    if ((S).vec == 4) {                                                    \
      switch ((S).g * 8 + (S).nch) {                                       \
        case 1 * 8 + 1: { constexpr int V = 4, G_ = 1, N_ = 1; CALL; } break;   \
        case 2 * 8 + 1: { constexpr int V = 4, G_ = 2, N_ = 1; CALL; } break;   \
        case 4 * 8 + 1: { constexpr int V = 4, G_ = 4, N_ = 1; CALL; } break;   \
        case 8 * 8 + 1: { constexpr int V = 4, G_ = 8, N_ = 1; CALL; } break;   \
        case 16 * 8 + 1: { constexpr int V = 4, G_ = 16, N_ = 1; CALL; } break; \
        case 32 * 8 + 1: { constexpr int V = 4, G_ = 32, N_ = 1; CALL; } break; \
        case 32 * 8 + 2: { constexpr int V = 4, G_ = 32, N_ = 2; CALL; } break; \
        case 32 * 8 + 4: { constexpr int V = 4, G_ = 32, N_ = 4; CALL; } break; \
        default: set_error("no kernel for shape"); return KGB_ERR_UNSUPPORTED;  \
      }                                                                    \
    } else {                                                               \
      switch ((S).g * 8 + (S).nch) {                                       \
        case 1 * 8 + 1: { constexpr int V = 1, G_ = 1, N_ = 1; CALL; } break;   \
        case 2 * 8 + 1: { constexpr int V = 1, G_ = 2, N_ = 1; CALL; } break;   \
        case 4 * 8 + 1: { constexpr int V = 1, G_ = 4, N_ = 1; CALL; } break;   \
        case 8 * 8 + 1: { constexpr int V = 1, G_ = 8, N_ = 1; CALL; } break;   \
        case 16 * 8 + 1: { constexpr int V = 1, G_ = 16, N_ = 1; CALL; } break; \
        case 32 * 8 + 1: { constexpr int V = 1, G_ = 32, N_ = 1; CALL; } break; \
        case 32 * 8 + 2: { constexpr int V = 1, G_ = 32, N_ = 2; CALL; } break; \
        case 32 * 8 + 4: { constexpr int V = 1, G_ = 32, N_ = 4; CALL; } break; \
        default: set_error("no kernel for shape"); return KGB_ERR_UNSUPPORTED;  \
      }                                                                    \
    }                                                                      \
  } while (0)

}  // namespace kgb

using namespace kgb;

extern "C" {

int32_t kgb_gather_unit_rows(void) { return kgb::UNIT_ROWS; }

size_t kgb_gather_reduce_partial_bytes(int32_t n_chunks, int32_t F, int32_t op) {
  if (n_chunks <= 0) return 0;
  size_t b = align_up((size_t)n_chunks * (size_t)F * sizeof(float), 256);
  if (op == KGB_OP_MAX || op == KGB_OP_MIN || op == KGB_OP_MAX_RAW) b *= 2;  // values + arg
  return b;
}

int kgb_gather_reduce(int device, const kgb_gather_reduce_args* a, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(a != nullptr, "args is NULL");
  KGB_REQUIRE(a->F > 0, "F must be positive (got %d)", a->F);
  KGB_REQUIRE(a->n_rows >= 0, "n_rows < 0");
  KGB_REQUIRE(a->op >= KGB_OP_SUM && a->op <= KGB_OP_SQDEV, "bad op %d", a->op);
  KGB_REQUIRE(a->op != KGB_OP_SQDEV || (a->addend && !a->bias && a->act == KGB_ACT_NONE && !a->row_ids),
              "KGB_OP_SQDEV needs the row means in `addend` and takes no bias / activation / row_ids");
  if (a->n_rows == 0) return KGB_OK;
  KGB_REQUIRE(a->x && a->rowptr && a->out, "x/rowptr/out must be non-NULL");
  KGB_REQUIRE(a->ldx >= a->F && a->ldo >= a->F, "leading dimension smaller than F");
  const bool hubs = a->n_hubs > 0 && a->n_chunks > 0;
  if (hubs) {
    KGB_REQUIRE(a->hub_row && a->hub_chunk_base && a->hub_nchunks && a->chunk_hub && a->partial,
                "hub table given but a pointer is NULL");
    KGB_REQUIRE(a->hub_chunk > 0 && a->hub_threshold >= a->hub_chunk, "bad hub threshold/chunk");
  }
  const bool is_max = (a->op == KGB_OP_MAX || a->op == KGB_OP_MIN || a->op == KGB_OP_MAX_RAW);
  cudaStream_t st = (cudaStream_t)stream;

  bool can4 = aligned16(a->x) && aligned16(a->out) && (a->ldx % 4 == 0) && (a->ldo % 4 == 0) &&
              (!a->addend || (aligned16(a->addend) && a->ld_addend % 4 == 0)) &&
              (!a->bias || aligned16(a->bias)) && (!a->arg || aligned16(a->arg)) &&
              (!hubs || aligned16(a->partial));
  // split source / split output (partitioned graphs)
  const bool split_out = a->out2 != nullptr || a->out2_push != nullptr;
  if (a->x2) {
    KGB_REQUIRE(a->n_split_src >= 0 && a->ldx2 >= a->F, "bad split source");
    can4 = can4 && aligned16(a->x2) && a->ldx2 % 4 == 0;
  }
  PushTab tab = {};
  if (split_out) {
    KGB_REQUIRE(!is_max && a->n_split_out >= 0, "a split output needs a linear aggregation");
    if (a->out2_push) {
      const kgb_halo_push_args* pa = a->out2_push;
      KGB_REQUIRE(pa->n_peers >= 1 && pa->n_peers <= KGB_MAX_PEERS && pa->ldd >= a->F, "bad push table");
      tab.n_peers = pa->n_peers;
      tab.ldd = pa->ldd;
      can4 = can4 && pa->ldd % 4 == 0;
      for (int q = 0; q <= KGB_MAX_PEERS; ++q) tab.slot_begin[q] = pa->slot_begin[q < pa->n_peers ? q : pa->n_peers];
      for (int q = 0; q < KGB_MAX_PEERS; ++q) {
        tab.dst[q] = q < pa->n_peers ? pa->dst[q] : nullptr;
        tab.dst_row0[q] = q < pa->n_peers ? pa->dst_row0[q] : 0;
        if (q < pa->n_peers && pa->slot_begin[q + 1] > pa->slot_begin[q]) {
          KGB_REQUIRE(pa->dst[q] != nullptr, "peer %d has rows but no window", q);
          can4 = can4 && aligned16(pa->dst[q]);
        }
      }
      KGB_REQUIRE(a->n_rows - a->n_split_out <= pa->slot_begin[pa->n_peers], "more split rows than push slots");
    } else {
      KGB_REQUIRE(a->ldo2 >= a->F, "bad split output");
      can4 = can4 && aligned16(a->out2) && a->ldo2 % 4 == 0;
    }
  }
  const int vec = (can4 && a->F % 4 == 0) ? 4 : 1;
  const int slab = 32 * 4 * vec;  // widest feature slab of one launch
  for (int f0 = 0; f0 < a->F; f0 += slab) {
    const int Fs = (a->F - f0 < slab) ? (a->F - f0) : slab;
    GRP p;
    p.x = a->x + f0; p.ldx = a->ldx; p.F = Fs;
    p.rowptr = a->rowptr; p.col = a->col; p.n_rows = a->n_rows; p.row_ids = a->row_ids;
    p.edge_w = a->edge_w; p.src_scale = a->src_scale; p.out_scale = a->out_scale;
    p.addend = a->addend ? a->addend + f0 : nullptr; p.ld_addend = a->ld_addend;
    p.addend_scale = a->addend_scale;
    p.bias = a->bias ? a->bias + f0 : nullptr; p.act = a->act;
    p.mean = (a->op == KGB_OP_MEAN); p.negate = (a->op == KGB_OP_MIN); p.raw_max = (a->op == KGB_OP_MAX_RAW);
    p.sqdev = (a->op == KGB_OP_SQDEV);
    p.out = a->out + f0; p.ldo = a->ldo; p.arg = a->arg ? a->arg + f0 : nullptr;
    p.hub_row = a->hub_row; p.hub_chunk_base = a->hub_chunk_base; p.hub_nchunks = a->hub_nchunks;
    p.chunk_hub = a->chunk_hub;
    p.n_hubs = hubs ? a->n_hubs : 0; p.n_chunks = hubs ? a->n_chunks : 0;
    p.hub_threshold = a->hub_threshold; p.hub_chunk = a->hub_chunk;
    p.partial = a->partial;
    p.work = a->work;
    p.unit_order = a->work ? a->unit_order : nullptr;
    p.x2 = a->x2 ? a->x2 + f0 : nullptr; p.ldx2 = a->ldx2;
    p.n_split_src = a->x2 ? a->n_split_src : INT64_MAX;
    p.out2 = a->out2 ? a->out2 + f0 : nullptr; p.ldo2 = a->ldo2;
    p.n_split_out = split_out ? a->n_split_out : INT64_MAX;
    p.col_hot = a->col_hot;
    p.edge_id = a->edge_id; p.drop_thr = 0u; p.drop_scale = 1.f;
    p.seed_lo = (uint32_t)(a->drop_seed & 0xffffffffull); p.seed_hi = (uint32_t)(a->drop_seed >> 32);
    if (a->drop_p > 0.f) {
      KGB_REQUIRE(a->drop_p < 1.f && a->edge_id != nullptr, "dropout needs 0 < p < 1 and the edge ids of the slots");
      p.drop_thr = dropout_threshold(a->drop_p);
      p.drop_scale = 1.f / (1.f - a->drop_p);
    }
    p.has_push = a->out2_push ? 1 : 0;
    p.tab = tab;
    if (p.has_push)
      for (int q = 0; q < tab.n_peers; ++q) if (p.tab.dst[q]) p.tab.dst[q] += f0;
    p.partial_arg = nullptr;
    if (hubs && is_max)
      p.partial_arg = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(a->partial) +
                                                 align_up((size_t)a->n_chunks * (size_t)a->F * sizeof(float), 256));
    // the partial buffer is reused by every slab (launches are stream-ordered)
    Shape s = pick_shape(Fs, vec == 4);
    int rc = KGB_OK;
    KGB_DISPATCH_SHAPE(s, (rc = launch_gr<V, G_, N_>(device, p, is_max, st)));
    if (rc != KGB_OK) return rc;
  }
  return KGB_OK;
}

size_t kgb_gather_max_bwd_workspace_bytes(int32_t n_hubs, int32_t n_chunks, int32_t F) {
  if (n_hubs <= 0 || n_chunks <= 0) return 0;
  return align_up((size_t)n_chunks * F * sizeof(float), 256) + align_up((size_t)n_hubs * F * sizeof(float), 256) +
         align_up((size_t)n_hubs * sizeof(int32_t), 256);
}

size_t kgb_gather_max_bwd_acc_bytes(int64_t n_src_rows, int32_t F) {
  if (n_src_rows <= 0 || F <= 0) return 0;
  return align_up((size_t)n_src_rows * (size_t)F * sizeof(long long), 256) + 256;
}

int kgb_gather_max_bwd(int device, const float* g, int64_t ldg, const int32_t* arg, const float* out,
                       int64_t ldo, const float* x, int64_t ldx, const int64_t* rowptr,
                       const int32_t* col, const int32_t* row_ids, int64_t n_rows, int32_t F,
                       int32_t op, float* gx, int64_t ldgx, int64_t n_src_rows, void* acc_ws,
                       const kgb_hub_table* hubs, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(op == KGB_OP_MAX || op == KGB_OP_MIN || op == KGB_OP_MAX_RAW, "op must be MAX or MIN");
  KGB_REQUIRE(F > 0 && n_rows >= 0 && n_src_rows >= 0, "bad sizes");
  if (n_rows == 0 || n_src_rows == 0) return KGB_OK;
  KGB_REQUIRE(g && arg && out && x && rowptr && col && gx && acc_ws, "NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  // fixed-point accumulators [n_src_rows, F] + the max|g| scalar behind them (zero on entry to the kernels)
  const size_t acc_bytes = kgb_gather_max_bwd_acc_bytes(n_src_rows, F);
  long long* acc = reinterpret_cast<long long*>(acc_ws);
  uint32_t* gmax_bits = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(acc_ws) + acc_bytes - 256);
  KGB_CHECK_CUDA(cudaMemsetAsync(acc_ws, 0, acc_bytes, st));
  {
    int64_t tg = ceil_div(n_rows, 8);
    const int64_t cap = (int64_t)sm_count(device) * 8;
    if (tg > cap) tg = cap;
    if (tg < 1) tg = 1;
    const int vec4 = (aligned16(g) && ldg % 4 == 0 && F % 4 == 0) ? 1 : 0;
    max_bwd_absmax_kernel<<<(int)tg, 256, 0, st>>>(g, ldg, n_rows, F, vec4, gmax_bits);
    KGB_CHECK_LAUNCH();
  }
  int cnt_bits = 1;
  while (((int64_t)1 << cnt_bits) < n_rows) ++cnt_bits;
  const bool use_hubs = hubs && hubs->n_hubs > 0 && hubs->n_chunks > 0 && hubs->partial;
  const bool can4 = aligned16(g) && aligned16(arg) && aligned16(out) && aligned16(x) && ldg % 4 == 0 &&
                    ldo % 4 == 0 && ldx % 4 == 0 && (!use_hubs || aligned16(hubs->partial));
  const int vec = (can4 && F % 4 == 0) ? 4 : 1;
  const int slab = 32 * 4 * vec;
  for (int f0 = 0; f0 < F; f0 += slab) {
    const int Fs = (F - f0 < slab) ? (F - f0) : slab;
    MaxBwdP p = {};
    p.g = g + f0; p.ldg = ldg; p.arg = arg + f0; p.out = out + f0; p.ldo = ldo; p.x = x + f0; p.ldx = ldx;
    p.rowptr = rowptr; p.col = col; p.row_ids = row_ids; p.n_rows = n_rows; p.F = Fs; p.gx = gx + f0; p.ldgx = ldgx;
    p.acc = acc + f0; p.ldacc = F; p.gmax_bits = gmax_bits; p.cnt_bits = cnt_bits;
    if (use_hubs) {
      p.hub_row = hubs->hub_row; p.hub_chunk_base = hubs->hub_chunk_base; p.hub_nchunks = hubs->hub_nchunks;
      p.chunk_hub = hubs->chunk_hub; p.n_hubs = hubs->n_hubs; p.n_chunks = hubs->n_chunks;
      p.hub_threshold = hubs->threshold; p.hub_chunk = hubs->chunk;
      char* w = reinterpret_cast<char*>(hubs->partial);
      p.cnt_partial = reinterpret_cast<float*>(w);
      w += align_up((size_t)p.n_chunks * F * sizeof(float), 256);
      p.cnt_total = reinterpret_cast<float*>(w);
      w += align_up((size_t)p.n_hubs * F * sizeof(float), 256);
      p.hub_tie = reinterpret_cast<int32_t*>(w);
      KGB_CHECK_CUDA(cudaMemsetAsync(p.hub_tie, 0, (size_t)p.n_hubs * sizeof(int32_t), st));
    }
    Shape s = pick_shape(Fs, vec == 4);
    const int grid = grid_for(device, n_rows, s.g);
    KGB_DISPATCH_SHAPE(s, (gather_max_bwd_kernel<V, G_, N_, 0><<<grid, 256, 0, st>>>(p)));
    KGB_CHECK_LAUNCH();
    if (use_hubs) {
      const int hgrid = grid_for(device, p.n_hubs, s.g);
      const int cgrid = grid_for(device, p.n_chunks, s.g);
      KGB_DISPATCH_SHAPE(s, (gather_max_bwd_kernel<V, G_, N_, 1><<<hgrid, 256, 0, st>>>(p)));
      KGB_CHECK_LAUNCH();
      KGB_DISPATCH_SHAPE(s, (max_bwd_hub_chunk_kernel<V, G_, N_, 0><<<cgrid, 256, 0, st>>>(p)));
      KGB_CHECK_LAUNCH();
      int64_t tg = ceil_div((int64_t)p.n_hubs * Fs, 256);
      if (tg > (int64_t)sm_count(device) * 8) tg = (int64_t)sm_count(device) * 8;
      max_bwd_hub_total_kernel<<<(int)tg, 256, 0, st>>>(p);
      KGB_CHECK_LAUNCH();
      KGB_DISPATCH_SHAPE(s, (max_bwd_hub_chunk_kernel<V, G_, N_, 1><<<cgrid, 256, 0, st>>>(p)));
      KGB_CHECK_LAUNCH();
    }
    {
      int64_t tg = ceil_div(n_src_rows, 8);
      const int64_t cap = (int64_t)sm_count(device) * 8;
      if (tg > cap) tg = cap;
      if (tg < 1) tg = 1;
      max_bwd_convert_kernel<<<(int)tg, 256, 0, st>>>(p, n_src_rows);
      KGB_CHECK_LAUNCH();
    }
  }
  return KGB_OK;
}

int kgb_gather_rows(int device, const float* src, int64_t lds, const int32_t* idx, int64_t n_out,
                    int32_t F, float scale, float* out, int64_t ldo, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(F > 0 && n_out >= 0, "bad sizes");
  if (n_out == 0) return KGB_OK;
  KGB_REQUIRE(src && out, "NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool can4 = aligned16(src) && aligned16(out) && lds % 4 == 0 && ldo % 4 == 0;
  const int vec = (can4 && F % 4 == 0) ? 4 : 1;
  const int slab = 32 * 4 * vec;
  for (int f0 = 0; f0 < F; f0 += slab) {
    const int Fs = (F - f0 < slab) ? (F - f0) : slab;
    Shape s = pick_shape(Fs, vec == 4);
    const int grid = grid_for(device, n_out, s.g);
    KGB_DISPATCH_SHAPE(s, (gather_rows_kernel<V, G_, N_><<<grid, 256, 0, st>>>(src + f0, lds, idx, n_out, Fs,
                                                                               scale, out + f0, ldo)));
    KGB_CHECK_LAUNCH();
  }
  return KGB_OK;
}

int kgb_halo_push(int device, const kgb_halo_push_args* a, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(a != nullptr, "args is NULL");
  KGB_REQUIRE(a->F > 0 && a->n_peers >= 1 && a->n_peers <= KGB_MAX_PEERS, "bad F / n_peers");
  KGB_REQUIRE(a->lds >= a->F && a->ldd >= a->F, "leading dimension smaller than F");
  PushTab tab;
  tab.n_peers = a->n_peers;
  tab.ldd = a->ldd;
  {
    const int64_t total = a->slot_begin[a->n_peers];
    tab.n_chunks = total > 0 ? ceil_div(total, (int64_t)PUSH_CHUNK) : 1;
    int64_t mul = (int64_t)((double)tab.n_chunks * 0.6180339887) | 1;   // golden-ratio stride, made coprime below
    auto gcd = [](int64_t x, int64_t y) { while (y) { const int64_t t = x % y; x = y; y = t; } return x; };
    while (mul > 1 && gcd(mul, tab.n_chunks) != 1) mul -= 2;
    if (mul < 1 || gcd(mul, tab.n_chunks) != 1) mul = 1;
    tab.perm_mul = mul;
    const int64_t rot = (a->slot_rot > 0 && a->slot_rot < total) ? a->slot_rot : 0;
    tab.perm_add = (rot / PUSH_CHUNK) % tab.n_chunks;
  }
  bool can4 = aligned16(a->src) && a->lds % 4 == 0 && a->ldd % 4 == 0;
  for (int p = 0; p <= a->n_peers; ++p) {
    tab.slot_begin[p] = a->slot_begin[p];
    KGB_REQUIRE(p == 0 ? a->slot_begin[0] == 0 : a->slot_begin[p] >= a->slot_begin[p - 1], "slot_begin must be a prefix sum");
  }
  for (int p = 0; p < KGB_MAX_PEERS; ++p) {
    tab.dst[p] = p < a->n_peers ? a->dst[p] : nullptr;
    tab.dst_row0[p] = p < a->n_peers ? a->dst_row0[p] : 0;
    if (p < a->n_peers && a->slot_begin[p + 1] > a->slot_begin[p]) {
      KGB_REQUIRE(a->dst[p] != nullptr && a->dst_row0[p] >= 0, "peer %d has slots but no window", p);
      can4 = can4 && aligned16(a->dst[p]);
    }
  }
  for (int p = a->n_peers + 1; p <= KGB_MAX_PEERS; ++p) tab.slot_begin[p] = a->slot_begin[a->n_peers];
  const int64_t n_slots = a->slot_begin[a->n_peers];
  if (n_slots == 0) return KGB_OK;
  KGB_REQUIRE(a->src != nullptr, "src is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = (can4 && a->F % 4 == 0) ? 4 : 1;
  const int slab = 32 * 4 * vec;
  for (int f0 = 0; f0 < a->F; f0 += slab) {
    const int Fs = (a->F - f0 < slab) ? (a->F - f0) : slab;
    PushTab t2 = tab;
    for (int p = 0; p < a->n_peers; ++p) if (t2.dst[p]) t2.dst[p] += f0;
    Shape s = pick_shape(Fs, vec == 4);
    const int grid = grid_for(device, n_slots, s.g);
    KGB_DISPATCH_SHAPE(s, (halo_push_kernel<V, G_, N_><<<grid, 256, 0, st>>>(a->src + f0, a->lds, a->idx, Fs, a->ldd, t2)));
    KGB_CHECK_LAUNCH();
  }
  return KGB_OK;
}

int kgb_dropout_mask(int device, const int32_t* edge_id, int64_t n_edges, int32_t F, int32_t per_head, float p,
                     uint64_t seed, float* out, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(n_edges >= 0 && F > 0 && p > 0.f && p < 1.f && out, "bad arguments");
  if (n_edges == 0) return KGB_OK;
  int64_t grid = ceil_div(n_edges * (int64_t)F, 256);
  const int64_t cap = (int64_t)sm_count(device) * 16;
  if (grid > cap) grid = cap;
  dropout_mask_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(edge_id, n_edges, F, per_head, dropout_threshold(p),
                                                                  1.f / (1.f - p), (uint32_t)(seed & 0xffffffffull),
                                                                  (uint32_t)(seed >> 32), out);
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

int kgb_relu_bwd(int device, const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows, int32_t F,
                 float* out, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(rows >= 0 && F > 0, "bad sizes");
  if (rows == 0) return KGB_OK;
  KGB_REQUIRE(g && y && out, "NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = rows * (int64_t)F;
  const int64_t cap = (int64_t)sm_count(device) * 16;
  if (ldg == F && ldy == F && total % 4 == 0 && aligned16(g) && aligned16(y) && aligned16(out)) {
    int64_t grid = ceil_div(total / 4, 256);
    if (grid > cap) grid = cap;
    relu_bwd_vec4_kernel<<<(int)grid, 256, 0, st>>>(reinterpret_cast<const float4*>(g),
                                                     reinterpret_cast<const float4*>(y), total / 4,
                                                     reinterpret_cast<float4*>(out));
  } else {
    int64_t grid = ceil_div(total, 256);
    if (grid > cap) grid = cap;
    relu_bwd_kernel<<<(int)grid, 256, 0, st>>>(g, ldg, y, ldy, rows, F, out);
  }
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

int kgb_permute_f32(int device, const float* in, const int32_t* perm, int64_t n, float* out,
                    kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  if (n <= 0) return KGB_OK;
  KGB_REQUIRE(in && perm && out, "NULL pointer");
  int64_t grid = ceil_div(n, 256);
  const int64_t cap = (int64_t)sm_count(device) * 16;
  if (grid > cap) grid = cap;
  permute_f32_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(in, perm, n, out);
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

int kgb_reduce_parts(int device, const float* part, int32_t n_parts, int32_t F, float* out,
                     kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(F > 0 && n_parts >= 0 && part && out, "bad arguments");
  reduce_parts_kernel<<<(F + 127) / 128, 128, 0, (cudaStream_t)stream>>>(part, n_parts, F, out);
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

}  // extern "C"
