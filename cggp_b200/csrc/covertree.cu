// Cover-tree inducing-point selection on the device - cggp/covertree.py:25-176.
//
// The reference builds the tree level by level with a greedy, order-dependent pass (first remaining row -> Lloyd mean
// of its neighbourhood -> the new node removes every row within `radius` from the neighbouring parents), a
// neighbour-list update and a Voronoi re-assignment of all rows per level.  Results identical to the reference need
// its order of operations AND NumPy's orders of summation; both are kept here:
//
//   * rows never move: every node owns an ordered list of ROW INDICES (int32) in HBM; "removing rows" is a stable
//     in-place compaction of such a list, "the rows of a node" a gather through it;
//   * greedy pass: ONE CTA runs the whole `while parent.data` loop of a parent (distance flags, block-wide scans,
//     ordered Lloyd mean, rejection test against the children of the neighbouring parents, stable compaction of every
//     neighbouring list) without returning to the host.  Two parents whose neighbour sets are disjoint touch disjoint
//     lists and children, so they commute exactly: the host orders the parents of a level into WAVES (a parent runs
//     one wave after the last earlier parent it shares a neighbour with) and each wave is one launch with one CTA per
//     parent.  At the fine levels that is hundreds of CTAs per launch; at the coarse levels (everything neighbours
//     everything) it degenerates to the reference's sequence, one CTA at a time;
//   * Voronoi pass: embarrassingly parallel - one thread per row takes the first minimum over the candidate children
//     of its parent; a STABLE radix sort of (child, row) then yields every child's list in exactly the reference's
//     order (parents in level order, rows in list order);
//   * summation orders (oracle/covertree.py spells them out and tests them against NumPy): row distance =
//     sqrt(pairwise sum over the features), mean of rows = row after row (pairwise for D = 1), mean of targets =
//     pairwise, the 1-D norm of the rejection test = a chain of fused multiply-adds.  No FMA contraction elsewhere.
//
// Entry points: cggp_covertree_build / _num_levels / _level_size / _level_radius / _level_points / _leaf_members /
// _cluster_stats / _destroy (include/cggp_b200.h).
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace ct {

constexpr int T = 512;      // threads of a greedy CTA
constexpr int NW = T / 32;
constexpr int ITEMS = 4;    // list entries per thread and chunk
constexpr int CHUNK = T * ITEMS;
constexpr int MAX_D = 128;  // one NumPy pairwise block: no recursion over the features
constexpr int VT = 256;     // threads of a Voronoi CTA
constexpr int VROWS = 1024; // rows per Voronoi work item

#ifdef __CUDA_ARCH__
#define CT_ADD(a, b) __dadd_rn((a), (b))
#define CT_SUB(a, b) __dsub_rn((a), (b))
#define CT_MUL(a, b) __dmul_rn((a), (b))
#else
#define CT_ADD(a, b) ((a) + (b))
#define CT_SUB(a, b) ((a) - (b))
#define CT_MUL(a, b) ((a) * (b))
#endif

// NumPy's pairwise summation of n <= 128 terms (one block: < 8 terms left to right, otherwise 8 accumulators).
template <class F>
__host__ __device__ __forceinline__ double pw_block(F term, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = CT_ADD(r, term(i));
    return r;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = term(j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = CT_ADD(r[j], term(i + j));
  }
  double res = CT_ADD(CT_ADD(CT_ADD(r[0], r[1]), CT_ADD(r[2], r[3])), CT_ADD(CT_ADD(r[4], r[5]), CT_ADD(r[6], r[7])));
  for (; i < n; ++i) res = CT_ADD(res, term(i));
  return res;
}

// np.linalg.norm(p - x, axis=-1) of one row
__host__ __device__ __forceinline__ double row_sq(const double* x, const double* p, int D) {
  return pw_block([&](int d) { const double t = CT_SUB(p[d], x[d]); return CT_MUL(t, t); }, D);
}
__device__ __forceinline__ double row_dist(const double* x, const double* p, int D) { return __dsqrt_rn(row_sq(x, p, D)); }

// NumPy's pairwise summation of a sequence of any length, one warp (all lanes get the sum): the 8 accumulators of a
// block are lanes 0..7.
template <class F>
__device__ double warp_pw_leaf(F value, int64_t lo, int n, int lane) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, value(lo + i));
    return r;
  }
  const int nb = n - (n % 8);
  double r = 0.0;
  if (lane < 8) {
    r = value(lo + lane);
    for (int i = 8; i < nb; i += 8) r = __dadd_rn(r, value(lo + i + lane));
  }
  r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
  r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
  r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
  double res = __shfl_sync(0xffffffffu, r, 0);
  for (int i = nb; i < n; ++i) res = __dadd_rn(res, value(lo + i));
  return res;
}
template <class F>
__device__ double warp_pw(F value, int64_t n, int lane) {
  struct Frame {
    int64_t lo, n;
    int stage;
    double left;
  };
  Frame st[48];
  int sp = 0;
  st[sp++] = {0, n, 0, 0.0};
  double ret = 0.0;
  while (sp > 0) {
    Frame& f = st[sp - 1];
    if (f.stage == 0) {
      if (f.n <= 128) {
        ret = warp_pw_leaf(value, f.lo, (int)f.n, lane);
        --sp;
      } else {
        int64_t n2 = f.n / 2;
        n2 -= n2 % 8;
        f.stage = 1;
        st[sp++] = {f.lo, n2, 0, 0.0};
      }
    } else if (f.stage == 1) {
      int64_t n2 = f.n / 2;
      n2 -= n2 % 8;
      f.left = ret;
      f.stage = 2;
      st[sp++] = {f.lo + n2, f.n - n2, 0, 0.0};
    } else {
      ret = __dadd_rn(f.left, ret);
      --sp;
    }
  }
  return ret;
}

// ------------------------------------------------------------------------------------------------------------------
// One pass of a CTA over an index list: flag = (distance of the row to `point` <= radius).
//   sel_out  != nullptr: the flagged entries are appended to sel_out (in order)
//   compact  == true   : the un-flagged entries are compacted to the front of the list (stable, in place)
// Returns the number of flagged entries (same value in every thread).
struct PassSmem {
  int cnt[ITEMS * NW];
  int off[ITEMS * NW];
  int total;
};

// Entries [lo, hi) of `src`: the kept (un-flagged) ones go to dst[kept_base ...], the flagged ones to
// sel_out[sel_base ...], both in list order.  dst == src with kept_base <= lo is a stable in-place compaction (every
// chunk is read completely before it is written, and writes never pass the read position).  COH: the lists are written
// by other CTAs of the cluster - read them through L2.  Returns the number of KEPT entries (same in every thread).
template <bool COH>
__device__ int64_t slice_pass(const double* __restrict__ X, int64_t ldx, int D, const int* src, int* dst, int64_t lo,
                              int64_t hi, int64_t kept_base, int64_t sel_base, const double* point, double radius,
                              int* sel_out, bool count_only, PassSmem& s) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int64_t kept_total = 0;
  int wcount = 0;
  // the entries of the NEXT chunk are fetched before the rows of this one are gathered: one memory latency per chunk
  // instead of two (in place this is safe: a chunk's writes never reach the positions of the next chunk)
  int nxt[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = lo + j * T + tid;
    nxt[j] = i < hi ? (COH ? __ldcg(src + i) : src[i]) : 0;
  }
  for (int64_t base = lo; base < hi; base += CHUNK) {
    int idx[ITEMS];
    bool keep[ITEMS], valid[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      valid[j] = base + j * T + tid < hi;
      idx[j] = nxt[j];
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int64_t i = base + CHUNK + j * T + tid;
      nxt[j] = i < hi ? (COH ? __ldcg(src + i) : src[i]) : 0;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      bool inside = false;
      if (valid[j]) inside = row_dist(X + (int64_t)idx[j] * ldx, point, D) <= radius;
      keep[j] = valid[j] && !inside;
      const unsigned m = __ballot_sync(0xffffffffu, keep[j]);
      if (count_only) wcount += __popc(m);
      else if (lane == 0) s.cnt[j * NW + w] = __popc(m);
    }
    if (count_only) continue;
    __syncthreads();
    if (w == 0) {
      const int v0 = s.cnt[2 * lane], v1 = s.cnt[2 * lane + 1];
      int incl = v0 + v1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int excl = incl - (v0 + v1);
      s.off[2 * lane] = excl;
      s.off[2 * lane + 1] = excl + v0;
      if (lane == 31) s.total = incl;
    }
    __syncthreads();
    const int kept_chunk = s.total;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const unsigned m = __ballot_sync(0xffffffffu, keep[j]);
      const int kr = s.off[j * NW + w] + __popc(m & ((1u << lane) - 1u));  // rank among the kept entries of the chunk
      if (keep[j]) {
        if (dst) dst[kept_base + kept_total + kr] = idx[j];
      } else if (valid[j] && sel_out) {
        const int pos = j * T + tid;  // position in the chunk, list order
        // flagged entries of the slice before this chunk + rank among the flagged entries of the chunk
        sel_out[sel_base + ((base - lo) - kept_total) + (pos - kr)] = idx[j];
      }
    }
    kept_total += kept_chunk;
  }
  if (count_only) {
    __syncthreads();  // protect s.cnt from a previous use
    if (lane == 0) s.cnt[w] = wcount;
    __syncthreads();
    for (int k = 0; k < NW; ++k) kept_total += s.cnt[k];
  }
  __syncthreads();
  return kept_total;
}

// One CTA over a whole list: flagged entries appended to sel_out, optional in-place compaction of the others.
// Returns the number of flagged entries.
__device__ int64_t list_pass(const double* __restrict__ X, int64_t ldx, int D, int* list, int64_t n, const double* point,
                             double radius, int* sel_out, bool compact, PassSmem& s) {
  return n - slice_pass<false>(X, ldx, D, list, compact ? list : nullptr, 0, n, 0, 0, point, radius, sel_out,
                               !compact && !sel_out, s);
}

// Ordered mean of the rows list[0..n) into point[0..D) (shared memory): x[list].mean(axis=-2).
template <bool COH = false>
__device__ void ordered_mean(const double* __restrict__ X, int64_t ldx, int D, const int* list, int64_t n, double* point,
                             double* stage, int stage_rows) {
  auto at = [&](int64_t i) { return COH ? __ldcg(list + i) : list[i]; };
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (D == 1) {
    if (w == 0) {
      const double s = warp_pw([&](int64_t i) { return X[(int64_t)at(i) * ldx]; }, n, lane);
      if (lane == 0) point[0] = __ddiv_rn(s, (double)n);
    }
    __syncthreads();
    return;
  }
  double acc[MAX_D / 32];
#pragma unroll
  for (int k = 0; k < MAX_D / 32; ++k) acc[k] = 0.0;
  int round = 0;
  for (int64_t r0 = 0; r0 < n; r0 += stage_rows, ++round) {
    double* buf = stage + (size_t)(round & 1) * stage_rows * D;
    const int rows = (int)min((int64_t)stage_rows, n - r0);
    if (w > 0) {
      for (int e = tid - 32; e < rows * D; e += T - 32) {
        const int row = e / D, d = e - row * D;
        buf[e] = X[(int64_t)at(r0 + row) * ldx + d];
      }
    }
    __syncthreads();
    if (w == 0) {
      int row = 0;
      if (r0 == 0) {  // NumPy starts from a copy of the first row
#pragma unroll
        for (int k = 0; k < MAX_D / 32; ++k)
          if (lane + 32 * k < D) acc[k] = buf[lane + 32 * k];
        row = 1;
      }
#pragma unroll 8
      for (; row < rows; ++row) {
#pragma unroll
        for (int k = 0; k < MAX_D / 32; ++k)
          if (lane + 32 * k < D) acc[k] = __dadd_rn(acc[k], buf[row * D + lane + 32 * k]);
      }
    }
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < MAX_D / 32; ++k)
      if (lane + 32 * k < D) point[lane + 32 * k] = __ddiv_rn(acc[k], (double)n);
  }
  __syncthreads();
}

struct GreedyArgs {
  const double* X;
  int64_t ldx;
  int D;
  double radius;
  int lloyds, store_children;
  const int* wave_parents;
  const int64_t* rn_off;
  const int* rn_idx;
  const int64_t* seg_off;
  int64_t* seg_cnt;
  int* data_idx;   // the lists, buffer 0
  int* data_alt;   // buffer 1 (cluster passes compact out of place and flip `which`)
  int* which;      // per parent: the buffer its list is in
  int* scratch;
  double* ch_pt;
  int* ch_next;
  int* p_first;
  int* p_last;
  int* n_children;
  int ch_cap;
  int64_t* ch_off;
  int64_t* ch_cnt;
  int* ch_pool;
  unsigned long long* pool_ptr;
  int* err;
  int stage_rows;
};

// covertree.py:68-101 for one parent per CTA
__global__ void __launch_bounds__(T) greedy_kernel(GreedyArgs a) {
  extern __shared__ double sm[];
  double* point = sm;
  double* init = sm + a.D;
  double* stage = sm + 2 * a.D;
  __shared__ PassSmem ps;
  __shared__ int s_child;
  __shared__ unsigned long long s_pool;
  const int tid = threadIdx.x, D = a.D;
  const int p = a.wave_parents[blockIdx.x];
  const int64_t off_p = a.seg_off[p];
  const int64_t rb = a.rn_off[p], re = a.rn_off[p + 1];
  volatile int64_t* seg_cnt = a.seg_cnt;
  volatile int* err = a.err;
  __shared__ int64_t s_cnt;
  auto list_of = [&](int r) { return (a.which[r] ? a.data_alt : a.data_idx) + a.seg_off[r]; };  // (no flips in this kernel)
  int* const list_p = list_of(p);
  while (true) {
    if (tid == 0) s_cnt = *err != 0 ? 0 : seg_cnt[p];  // another CTA's failure stops everybody
    __syncthreads();
    const int64_t cnt = s_cnt;
    if (cnt == 0) break;
    const int first = list_p[0];
    if (tid < D) {
      const double v = a.X[(int64_t)first * a.ldx + tid];
      init[tid] = v;
      point[tid] = v;
    }
    __syncthreads();
    if (a.lloyds) {
      int* sel = a.scratch + off_p;
      const int64_t nsel = list_pass(a.X, a.ldx, D, list_p, cnt, init, a.radius, sel, false, ps);
      // nsel >= 1: the first row is at distance 0 from itself
      ordered_mean(a.X, a.ldx, D, sel, nsel, point, stage, a.stage_rows);
      // rejected when it comes within `radius` of an existing child of a neighbouring parent (covertree.py:76-83)
      int hit = 0;
      for (int64_t ri = rb + tid; ri < re; ri += T) {
        const int r = a.rn_idx[ri];
        for (int c = __ldcg(a.p_first + r); c >= 0 && !hit; c = __ldcg(a.ch_next + c)) {
          double s = 0.0;
          for (int d = 0; d < D; ++d) {
            const double t = __dsub_rn(point[d], __ldcg(a.ch_pt + (int64_t)c * D + d));
            s = __fma_rn(t, t, s);
          }
          if (__dsqrt_rn(s) < a.radius) hit = 1;
        }
      }
      if (__syncthreads_or(hit)) {
        if (tid < D) point[tid] = init[tid];
        __syncthreads();
      }
    }
    // the new child (covertree.py:99-101)
    if (tid == 0) {
      const int id = atomicAdd(a.n_children, 1);
      if (id >= a.ch_cap) {
        *err = 2;
        s_child = -1;
      } else {
        a.ch_next[id] = -1;
        const int last = a.p_last[p];
        if (last < 0) a.p_first[p] = id; else a.ch_next[last] = id;
        a.p_last[p] = id;
        s_child = id;
      }
    }
    __syncthreads();
    const int child = s_child;
    if (child < 0) break;
    if (tid < D) a.ch_pt[(int64_t)child * D + tid] = point[tid];
    int* pool = nullptr;
    if (a.store_children) {
      int64_t total = 0;
      for (int64_t ri = rb; ri < re; ++ri) {
        const int r = a.rn_idx[ri];
        const int64_t n = seg_cnt[r];
        if (n > 0) total += list_pass(a.X, a.ldx, D, list_of(r), n, point, a.radius, nullptr, false, ps);
      }
      if (tid == 0) {
        s_pool = atomicAdd(a.pool_ptr, (unsigned long long)total);
        a.ch_off[child] = (int64_t)s_pool;
        a.ch_cnt[child] = total;
      }
      __syncthreads();
      pool = a.ch_pool + s_pool;
    }
    // every row within `radius` leaves the neighbouring parents' lists (covertree.py:86-97)
    int64_t taken_here = 0;
    for (int64_t ri = rb; ri < re; ++ri) {
      const int r = a.rn_idx[ri];
      const int64_t n = seg_cnt[r];
      if (n == 0) continue;
      const int64_t taken = list_pass(a.X, a.ldx, D, list_of(r), n, point, a.radius, pool, true, ps);
      if (pool) pool += taken;
      if (r == p) taken_here = taken;
      if (tid == 0 && taken) seg_cnt[r] = n - taken;
    }
    __syncthreads();
    if (taken_here == 0) {  // the reference would loop forever here
      if (tid == 0) *err = 1;
      break;
    }
  }
}

// The same loop for ONE parent per thread-block CLUSTER (2, 4, 8 or 16 CTAs), used for the waves with few parents, where one
// CTA per parent would leave most SMs idle.  Two ways of sharing a step:
//   * by slices (few, long lists - the coarse levels, where a step walks up to N rows): every CTA counts the kept
//     entries of its contiguous slice -> the counts are exchanged through distributed shared memory (one cluster
//     barrier) -> every CTA writes its slice at its prefix, kept entries into the OTHER buffer of the list (`which`
//     flips), flagged entries to the selection / the new node's rows;
//   * by lists (>= 2 lists per CTA - the fine levels): the CTAs take the neighbouring lists in turn and compact them in
//     place, one cluster barrier per step.
// The ordered Lloyd mean, the rejection test and the node table stay with CTA 0, which broadcasts the outcome into the
// other CTAs' shared memory.
struct ClusterShared {
  int64_t kept[2][16];  // per pass parity: kept entries of every CTA's slice
  int64_t cnt;         // rows left in the parent (0 = stop), broadcast by CTA 0
  int child;
  unsigned long long pool;
};

__global__ void __launch_bounds__(T) greedy_cluster_kernel(GreedyArgs a) {
  namespace cgx = cooperative_groups;
  cgx::cluster_group cluster = cgx::this_cluster();
  const int rank = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  extern __shared__ double sm[];
  double* point = sm;
  double* init = sm + a.D;
  double* stage = sm + 2 * a.D;
  __shared__ PassSmem ps;
  __shared__ ClusterShared cs;
  const int tid = threadIdx.x, D = a.D;
  const int p = a.wave_parents[blockIdx.x / CS];
  const int64_t off_p = a.seg_off[p];
  const int64_t rb = a.rn_off[p], re = a.rn_off[p + 1];
  volatile int64_t* seg_cnt = a.seg_cnt;
  volatile int* which = a.which;
  volatile int* err = a.err;
  int pass_no = 0;

  // cooperative pass over the list of node r (n rows): returns the number of flagged rows
  auto coop_pass = [&](int r, int64_t n, const double* pt, int* sel_out, bool compact) -> int64_t {
    const int wh = which[r];
    const int* src = (wh ? a.data_alt : a.data_idx) + a.seg_off[r];
    int* dst = (wh ? a.data_idx : a.data_alt) + a.seg_off[r];
    const int64_t nch = (n + CHUNK - 1) / CHUNK, per = (nch + CS - 1) / CS;
    const int64_t lo = min(n, (int64_t)rank * per * CHUNK), hi = min(n, (int64_t)(rank + 1) * per * CHUNK);
    const int64_t kept_local = slice_pass<true>(a.X, a.ldx, D, src, nullptr, lo, hi, 0, 0, pt, a.radius, nullptr, true, ps);
    const int par = pass_no & 1;
    ++pass_no;
    if (tid < CS) *cluster.map_shared_rank(&cs.kept[par][rank], tid) = kept_local;
    cluster.sync();
    int64_t prefix = 0, total = 0;
    for (int c = 0; c < CS; ++c) {
      const int64_t v = cs.kept[par][c];
      if (c < rank) prefix += v;
      total += v;
    }
    const int64_t flagged = n - total;
    if (flagged > 0 && (sel_out || compact)) {
      slice_pass<true>(a.X, a.ldx, D, src, compact ? dst : nullptr, lo, hi, prefix, lo - prefix, pt, a.radius, sel_out,
                       false, ps);
      if (compact && rank == 0 && tid == 0) {
        which[r] = wh ^ 1;
        seg_cnt[r] = total;
      }
    }
    return flagged;
  };

  while (true) {
    if (rank == 0 && tid == 0) {
      const int64_t v = *err != 0 ? 0 : seg_cnt[p];
      for (int c = 0; c < CS; ++c) *cluster.map_shared_rank(&cs.cnt, c) = v;
    }
    cluster.sync();
    const int64_t cnt = cs.cnt;
    if (cnt == 0) break;
    const int first = __ldcg((which[p] ? a.data_alt : a.data_idx) + off_p);
    if (tid < D) {
      const double v = a.X[(int64_t)first * a.ldx + tid];
      init[tid] = v;
      point[tid] = v;
    }
    __syncthreads();
    if (a.lloyds) {
      int* sel = a.scratch + off_p;
      const int64_t nsel = coop_pass(p, cnt, init, sel, false);
      cluster.sync();  // the selection is complete
      if (rank == 0) {
        ordered_mean<true>(a.X, a.ldx, D, sel, nsel, point, stage, a.stage_rows);
        int hit = 0;
        for (int64_t ri = rb + tid; ri < re; ri += T) {
          const int r = a.rn_idx[ri];
          for (int c = __ldcg(a.p_first + r); c >= 0 && !hit; c = __ldcg(a.ch_next + c)) {
            double s = 0.0;
            for (int d = 0; d < D; ++d) {
              const double t = __dsub_rn(point[d], __ldcg(a.ch_pt + (int64_t)c * D + d));
              s = __fma_rn(t, t, s);
            }
            if (__dsqrt_rn(s) < a.radius) hit = 1;
          }
        }
        if (__syncthreads_or(hit)) {
          if (tid < D) point[tid] = init[tid];
          __syncthreads();
        }
      }
    }
    // the new child: CTA 0 enters it into the node table and tells the others
    if (rank == 0) {
      if (tid == 0) {
        const int id = atomicAdd(a.n_children, 1);
        int child = -1;
        if (id >= a.ch_cap) {
          *err = 2;
        } else {
          a.ch_next[id] = -1;
          const int last = a.p_last[p];
          if (last < 0) a.p_first[p] = id; else a.ch_next[last] = id;
          a.p_last[p] = id;
          child = id;
        }
        for (int c = 0; c < CS; ++c) *cluster.map_shared_rank(&cs.child, c) = child;
      }
      if (tid < D) {
        const double v = point[tid];
        for (int c = 1; c < CS; ++c) cluster.map_shared_rank(point, c)[tid] = v;
      }
    }
    cluster.sync();
    const int child = cs.child;
    if (child < 0) break;
    if (rank == 0 && tid < D) a.ch_pt[(int64_t)child * D + tid] = point[tid];
    int* pool = nullptr;
    if (a.store_children) {
      int64_t total = 0;
      for (int64_t ri = rb; ri < re; ++ri) {
        const int r = a.rn_idx[ri];
        const int64_t n = seg_cnt[r];
        if (n > 0) total += coop_pass(r, n, point, nullptr, false);
      }
      if (rank == 0 && tid == 0) {
        const unsigned long long base = atomicAdd(a.pool_ptr, (unsigned long long)total);
        a.ch_off[child] = (int64_t)base;
        a.ch_cnt[child] = total;
        for (int c = 0; c < CS; ++c) *cluster.map_shared_rank(&cs.pool, c) = base;
      }
      cluster.sync();
      pool = a.ch_pool + cs.pool;
    }
    if (!a.store_children && re - rb >= 2 * CS) {
      // many neighbouring lists: the CTAs of the cluster take them in turn, each compacting its lists in place
      for (int64_t ri = rb + rank; ri < re; ri += CS) {
        const int r = a.rn_idx[ri];
        const int64_t n = seg_cnt[r];
        if (n == 0) continue;
        int* list = (which[r] ? a.data_alt : a.data_idx) + a.seg_off[r];
        const int64_t kept = slice_pass<true>(a.X, a.ldx, D, list, list, 0, n, 0, 0, point, a.radius, nullptr, false, ps);
        if (tid == 0) {
          if (kept != n) seg_cnt[r] = kept;
          if (r == p && kept == n) *err = 1;  // the reference would loop forever here
        }
      }
      cluster.sync();  // all lists and counts are in place before CTA 0 looks at the parent again
      continue;
    }
    int64_t taken_here = 0;
    for (int64_t ri = rb; ri < re; ++ri) {
      const int r = a.rn_idx[ri];
      const int64_t n = seg_cnt[r];
      if (n == 0) continue;
      const int64_t taken = coop_pass(r, n, point, pool, true);
      if (pool) pool += taken;
      if (r == p) taken_here = taken;
    }
    if (taken_here == 0) {  // the reference would loop forever here
      if (rank == 0 && tid == 0) *err = 1;
      break;
    }
  }
  cluster.sync();  // nobody leaves while its shared memory may still be written
}

// root: mean of all rows (covertree.py:49)
__global__ void __launch_bounds__(T) root_mean_kernel(const double* X, int64_t ldx, int D, const int* list, int64_t n,
                                                      double* out, int stage_rows) {
  extern __shared__ double sm[];
  ordered_mean(X, ldx, D, list, n, sm, sm + D, stage_rows);
  if (threadIdx.x < D) out[threadIdx.x] = sm[threadIdx.x];
}
__global__ void iota_kernel(int* v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = (int)i;
}
// largest distance of a row to the root mean (covertree.py:50-51); distances are >= 0, so their bit patterns order
__global__ void __launch_bounds__(256) max_dist_kernel(const double* X, int64_t ldx, int D, int64_t n, const double* mean,
                                                       unsigned long long* out) {
  __shared__ double pm[MAX_D];
  __shared__ unsigned long long red[8];
  if (threadIdx.x < D) pm[threadIdx.x] = mean[threadIdx.x];
  __syncthreads();
  unsigned long long best = 0ull;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(row_dist(X + i * ldx, pm, D));
    best = max(best, b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) best = max(best, red[k]);
    atomicMax(out, best);
  }
}

// Voronoi pass (covertree.py:117-155): key of a row = the nearest candidate child of its parent, first minimum.
// sqrt is monotone, so a candidate whose SQUARED distance is not smaller cannot win; the others are compared exactly
// as NumPy does, on the rounded square roots.
__global__ void __launch_bounds__(VT) voronoi_kernel(const double* __restrict__ X, int64_t ldx, int D,
                                                     const int* __restrict__ vor, const int64_t* __restrict__ desc,
                                                     const int64_t* __restrict__ cand_off,
                                                     const int* __restrict__ cand_idx,
                                                     const double* __restrict__ child_pts, int* __restrict__ keys,
                                                     int smem_doubles) {
  extern __shared__ double cp[];
  const int64_t p = desc[3 * blockIdx.x], start = desc[3 * blockIdx.x + 1], len = desc[3 * blockIdx.x + 2];
  const int64_t cb = cand_off[p];
  const int nc = (int)(cand_off[p + 1] - cb);
  const bool staged = (int64_t)nc * D <= smem_doubles;
  if (staged) {
    for (int e = threadIdx.x; e < nc * D; e += VT) {
      const int c = e / D, d = e - c * D;
      cp[e] = child_pts[(int64_t)cand_idx[cb + c] * D + d];
    }
  }
  __syncthreads();
  for (int64_t i = start + threadIdx.x; i < start + len; i += VT) {
    const double* x = X + (int64_t)vor[i] * ldx;
    double best_s = 0.0, best_r = 0.0;
    int best = -1;
    for (int c = 0; c < nc; ++c) {
      const double* q = staged ? cp + c * D : child_pts + (int64_t)cand_idx[cb + c] * D;
      const double s = row_sq(x, q, D);
      if (best < 0) {
        best = c; best_s = s; best_r = __dsqrt_rn(s);
      } else if (s < best_s) {
        const double r = __dsqrt_rn(s);
        if (r < best_r) { best = c; best_s = s; best_r = r; }
      }
    }
    keys[i] = cand_idx[cb + best];
  }
}

// offsets of the runs of a sorted key array: off[k] = first position with key >= k, off[nk] = n
__global__ void bounds_kernel(const int* __restrict__ keys, int64_t n, int nk, int64_t* __restrict__ off) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
    const int kprev = i == 0 ? -1 : keys[i - 1];
    const int kcur = i == n ? nk : keys[i];
    for (int k = kprev + 1; k <= kcur; ++k) off[k] = i;
  }
}
__global__ void counts_kernel(const int64_t* __restrict__ off, int nk, int64_t* __restrict__ cnt) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nk; k += gridDim.x * blockDim.x) cnt[k] = off[k + 1] - off[k];
}

// np.mean(node.data[1]) and the count per leaf (covertree.py:168-176): one warp per leaf, pairwise summation
template <typename T_>
__global__ void __launch_bounds__(256) leaf_stats_kernel(const T_* __restrict__ y, int64_t ldy, const int* __restrict__ idx,
                                                         const int64_t* __restrict__ off, const int64_t* __restrict__ cnt,
                                                         int64_t m, T_* __restrict__ means, T_* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t leaf = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (leaf >= m) return;
  const int64_t n = cnt[leaf];
  const int* list = idx + off[leaf];
  const double s = warp_pw([&](int64_t i) { return (double)y[(int64_t)list[i] * ldy]; }, n, lane);
  if (lane == 0) {
    means[leaf] = (T_)__ddiv_rn(s, (double)n);  // empty leaf: 0 / 0 = nan, as np.mean of nothing
    counts[leaf] = (T_)n;
  }
}
__global__ void members_kernel(const int* __restrict__ idx, const int64_t* __restrict__ off,
                               const int64_t* __restrict__ cnt, int64_t m, int64_t n, int64_t* __restrict__ out_off,
                               int64_t* __restrict__ out_rows) {
  // leaf lists are stored back to back in leaf order (Voronoi mode) or scattered (pool mode): copy list by list
  const int64_t leaf = blockIdx.x;
  const int64_t c = cnt[leaf], o = off[leaf], dst = out_off[leaf];
  for (int64_t i = threadIdx.x; i < c; i += blockDim.x) out_rows[dst + i] = idx[o + i];
  (void)m; (void)n;
}
__global__ void exclusive_offsets_kernel(const int64_t* __restrict__ cnt, int64_t m, int64_t* __restrict__ out_off) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t run = 0;
    for (int64_t k = 0; k < m; ++k) { out_off[k] = run; run += cnt[k]; }
    out_off[m] = run;
  }
}

// Bump allocator over the ctx workspace (kept between calls: cudaMalloc / cudaFree of hundreds of MB per build cost
// up to 270 ms) - `base == nullptr` only measures.
struct Arena {
  char* base = nullptr;
  size_t used = 0;
  void* take(size_t bytes) {
    used = (used + 255) & ~(size_t)255;
    void* p = base ? base + used : nullptr;
    used += bytes;
    return p;
  }
};
template <typename U>
struct DevBuf {
  U* p = nullptr;
  size_t n = 0;
  bool owned = false;
  ~DevBuf() { if (p && owned) cudaFree(p); }
  cudaError_t alloc(size_t count) {  // own allocation (small per-level tables)
    if (p && owned) cudaFree(p);
    p = nullptr;
    n = count;
    owned = true;
    return cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(U));
  }
  void carve(Arena& a, size_t count) {  // a piece of the workspace
    n = count;
    owned = false;
    p = (U*)a.take(std::max<size_t>(count, 1) * sizeof(U));
  }
};

struct Level {
  int64_t size = 0;
  double radius = 0.0;
  std::vector<double> pts;                 // [size, D]
  std::vector<int> parent;                 // index in the previous level
  std::vector<std::vector<int>> rn;        // r_neighbors (indices in this level, reference order)
  std::vector<int> child_begin, child_end; // children are contiguous in the next level
};

}  // namespace ct

struct cggp_covertree {
  int D = 0;
  int64_t n = 0;
  int device = 0;
  std::vector<ct::Level> levels;
  // rows of the last level's nodes
  int* leaf_idx = nullptr;
  int64_t* leaf_off = nullptr;
  int64_t* leaf_cnt = nullptr;
  ~cggp_covertree() {
    if (leaf_idx) cudaFree(leaf_idx);
    if (leaf_off) cudaFree(leaf_off);
    if (leaf_cnt) cudaFree(leaf_cnt);
  }
};

using namespace ct;

extern "C" int cggp_covertree_build(cggp_ctx* ctx, int dtype, const void* dev_X, int64_t n, int D, int64_t ldx,
                                    double spatial_resolution, int num_levels, int lloyds, int voronoi,
                                    cggp_covertree** out) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (!out) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "covertree: null output");
  *out = nullptr;
  if (dtype != CGGP_F64) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "covertree: float64 rows only");
  if (!dev_X || n <= 0 || n >= 0x7fffffff || D <= 0 || ldx < D) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "covertree: bad arguments");
  if (D > MAX_D) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "covertree: D > %d", MAX_D);
  const double* X = (const double*)dev_X;
  cudaStream_t st = ctx->stream;

  // shared memory of the greedy kernel: point, first row, two stages of rows for the ordered mean
  int stage_rows = std::max(8, std::min(2048, (int)((96 * 1024) / (sizeof(double) * 2 * D))));
  const size_t greedy_smem = sizeof(double) * (2 * (size_t)D + 2 * (size_t)stage_rows * D);
  CGGP_CUDA(ctx, cudaFuncSetAttribute(greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)greedy_smem));
  CGGP_CUDA(ctx, cudaFuncSetAttribute(greedy_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)greedy_smem));
  CGGP_CUDA(ctx, cudaFuncSetAttribute(root_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)greedy_smem));
  // waves with few parents run one parent per cluster of CTAs (CGGP_CT_CLUSTER = 0: never, 1: always, default: where a
  // parent holds >= 32 rows on average; results do not depend on it)
  const char* cl_env = getenv("CGGP_CT_CLUSTER");
  const int cl_mode = cl_env ? atoi(cl_env) : -1;
  const bool trace = getenv("CGGP_CT_TRACE") != nullptr;
  // clusters of 16 CTAs (non-portable size) where the device schedules them, 8 otherwise
  int max_cluster = 8;
  if (cudaFuncSetAttribute(greedy_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = greedy_smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, greedy_cluster_kernel, &cfg) == cudaSuccess && nclusters >= 1)
      max_cluster = 16;
  }
  (void)cudaGetLastError();
  const auto t_begin = std::chrono::steady_clock::now();
  auto ms_since = [](std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  };
  const int vor_smem_doubles = (96 * 1024) / sizeof(double);
  CGGP_CUDA(ctx, cudaFuncSetAttribute(voronoi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(vor_smem_doubles * sizeof(double))));

  DevBuf<int> data_idx, data_alt, which, scratch, vorA, vorB, keysA, keysB, ch_next, p_first, p_last, pool, counters;
  DevBuf<double> ch_pt, small;
  DevBuf<int64_t> seg_off, seg_cnt, ch_off, ch_cnt;
  DevBuf<unsigned long long> ull;
  DevBuf<unsigned char> cub_tmp;
  size_t cub_bytes = 0;
  if (voronoi)
    CGGP_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int*)nullptr, (int*)nullptr,
                                                   (const int*)nullptr, (int*)nullptr, n, 0, 31, st));
  auto layout = [&](Arena& ar) {
    data_idx.carve(ar, n);
    data_alt.carve(ar, n);
    which.carve(ar, n);
    scratch.carve(ar, n);
    ch_next.carve(ar, n);
    p_first.carve(ar, n);
    p_last.carve(ar, n);
    ch_pt.carve(ar, (size_t)n * D);
    seg_off.carve(ar, n + 1);
    seg_cnt.carve(ar, n + 1);
    counters.carve(ar, 4);  // [0] number of children, [1] error flag
    ull.carve(ar, 2);       // [0] pool pointer, [1] largest root distance
    small.carve(ar, MAX_D);
    if (voronoi) {
      vorA.carve(ar, n);
      vorB.carve(ar, n);
      keysA.carve(ar, n);
      keysB.carve(ar, n);
      cub_tmp.carve(ar, cub_bytes);
    } else {
      pool.carve(ar, n);
      ch_off.carve(ar, n);
      ch_cnt.carve(ar, n);
    }
  };
  {
    Arena measure;
    layout(measure);
    const int rc = cggp_ws_reserve(ctx, measure.used + 256);
    if (rc) return rc;
    Arena real;
    real.base = (char*)ctx->ws;
    layout(real);
  }
  const int iota_grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);

  // ---- root (covertree.py:49-63) ----
  iota_kernel<<<iota_grid, 256, 0, st>>>(data_idx.p, n);
  CGGP_LAUNCH_CHECK(ctx);
  root_mean_kernel<<<1, T, greedy_smem, st>>>(X, ldx, D, data_idx.p, n, small.p, stage_rows);
  CGGP_LAUNCH_CHECK(ctx);
  CGGP_CUDA(ctx, cudaMemsetAsync(ull.p, 0, 2 * sizeof(unsigned long long), st));
  max_dist_kernel<<<iota_grid, 256, 0, st>>>(X, ldx, D, n, small.p, ull.p + 1);
  CGGP_LAUNCH_CHECK(ctx);
  std::vector<double> root_mean(D);
  unsigned long long maxbits = 0;
  CGGP_CUDA(ctx, cudaMemcpyAsync(root_mean.data(), small.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
  CGGP_CUDA(ctx, cudaMemcpyAsync(&maxbits, ull.p + 1, sizeof(maxbits), cudaMemcpyDeviceToHost, st));
  CGGP_CUDA(ctx, cudaStreamSynchronize(st));
  double max_radius;
  memcpy(&max_radius, &maxbits, sizeof(double));
  if (spatial_resolution > 0.0) {
    // (all rows equal: the reference's math.log2(0) raises as well)
    if (!(max_radius > 0.0)) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "covertree: all rows coincide, no level count for a resolution");
    num_levels = (int)std::ceil(std::log2(max_radius / spatial_resolution)) + 1;
    if (num_levels >= 1) max_radius = spatial_resolution * (double)(1ll << (num_levels - 1));
  }
  if (num_levels < 1 || num_levels > 62)
    CGGP_FAIL(ctx, CGGP_ERR_INVALID, "covertree: %d levels (largest distance to the mean %g, resolution %g)", num_levels,
              max_radius, spatial_resolution);

  cggp_covertree* tree = new cggp_covertree();
  struct TreeGuard {
    cggp_covertree*& t;
    bool keep = false;
    ~TreeGuard() { if (!keep) { delete t; t = nullptr; } }
  } guard{tree};
  tree->D = D;
  tree->n = n;
  tree->device = ctx->device;
  tree->levels.resize(num_levels);
  if (trace) fprintf(stderr, "covertree: buffers + root (mean, largest distance) %.3f ms\n", ms_since(t_begin));
  {
    Level& root = tree->levels[0];
    root.size = 1;
    root.radius = max_radius;
    root.pts = root_mean;
    root.parent = {-1};
    root.rn = {{0}};
  }
  // rows of the root: all of them, in order
  int64_t one_off = 0, one_cnt = n;
  CGGP_CUDA(ctx, cudaMemcpyAsync(seg_off.p, &one_off, sizeof(int64_t), cudaMemcpyHostToDevice, st));
  CGGP_CUDA(ctx, cudaMemcpyAsync(seg_cnt.p, &one_cnt, sizeof(int64_t), cudaMemcpyHostToDevice, st));
  int* vor_cur = vorA.p;
  int* vor_alt = vorB.p;
  if (voronoi) CGGP_CUDA(ctx, cudaMemcpyAsync(vor_cur, data_idx.p, sizeof(int) * n, cudaMemcpyDeviceToDevice, st));
  int* data_cur = data_idx.p;
  int* pool_cur = pool.p;

  for (int level = 1; level < num_levels; ++level) {
    Level& par = tree->levels[level - 1];
    Level& lv = tree->levels[level];
    const int64_t P = par.size;
    const double radius = max_radius / std::ldexp(1.0, level);
    lv.radius = radius;
    const double nfactor = 4.0 * (1.0 - 1.0 / std::ldexp(1.0, num_levels - level));  // neighbor_factor[level], :64

    // neighbour lists of the parents as CSR + the wave schedule
    std::vector<int64_t> rn_off(P + 1, 0);
    for (int64_t i = 0; i < P; ++i) rn_off[i + 1] = rn_off[i] + (int64_t)par.rn[i].size();
    std::vector<int> rn_idx((size_t)rn_off[P]);
    for (int64_t i = 0; i < P; ++i) std::copy(par.rn[i].begin(), par.rn[i].end(), rn_idx.begin() + rn_off[i]);
    std::vector<int> last_wave(P, 0), wave(P, 0);
    int n_waves = 0;
    for (int64_t i = 0; i < P; ++i) {
      int wv = 0;
      for (int q : par.rn[i]) wv = std::max(wv, last_wave[q]);
      wave[i] = wv + 1;
      for (int q : par.rn[i]) last_wave[q] = wave[i];
      n_waves = std::max(n_waves, wave[i]);
    }
    std::vector<int64_t> wave_off(n_waves + 2, 0);
    for (int64_t i = 0; i < P; ++i) wave_off[wave[i] + 1]++;
    for (int k = 1; k <= n_waves + 1; ++k) wave_off[k] += wave_off[k - 1];
    std::vector<int> wave_parents(P);
    {
      std::vector<int64_t> fill(wave_off.begin(), wave_off.end());
      for (int64_t i = 0; i < P; ++i) wave_parents[fill[wave[i]]++] = (int)i;
    }
    DevBuf<int64_t> d_rn_off;
    DevBuf<int> d_rn_idx, d_wave_parents;
    CGGP_CUDA(ctx, d_rn_off.alloc(P + 1));
    CGGP_CUDA(ctx, d_rn_idx.alloc(rn_idx.size()));
    CGGP_CUDA(ctx, d_wave_parents.alloc(P));
    CGGP_CUDA(ctx, cudaMemcpyAsync(d_rn_off.p, rn_off.data(), sizeof(int64_t) * (P + 1), cudaMemcpyHostToDevice, st));
    if (!rn_idx.empty())
      CGGP_CUDA(ctx, cudaMemcpyAsync(d_rn_idx.p, rn_idx.data(), sizeof(int) * rn_idx.size(), cudaMemcpyHostToDevice, st));
    CGGP_CUDA(ctx, cudaMemcpyAsync(d_wave_parents.p, wave_parents.data(), sizeof(int) * P, cudaMemcpyHostToDevice, st));
    CGGP_CUDA(ctx, cudaMemsetAsync(counters.p, 0, 4 * sizeof(int), st));
    CGGP_CUDA(ctx, cudaMemsetAsync(ull.p, 0, sizeof(unsigned long long), st));
    CGGP_CUDA(ctx, cudaMemsetAsync(p_first.p, 0xff, sizeof(int) * P, st));
    CGGP_CUDA(ctx, cudaMemsetAsync(p_last.p, 0xff, sizeof(int) * P, st));
    CGGP_CUDA(ctx, cudaMemsetAsync(which.p, 0, sizeof(int) * P, st));
    const auto t_level = std::chrono::steady_clock::now();

    GreedyArgs ga;
    ga.X = X; ga.ldx = ldx; ga.D = D; ga.radius = radius; ga.lloyds = lloyds; ga.store_children = voronoi ? 0 : 1;
    ga.rn_off = d_rn_off.p; ga.rn_idx = d_rn_idx.p; ga.seg_off = seg_off.p; ga.seg_cnt = seg_cnt.p;
    ga.data_idx = data_cur; ga.data_alt = data_alt.p; ga.which = which.p; ga.scratch = scratch.p; ga.ch_pt = ch_pt.p; ga.ch_next = ch_next.p;
    ga.p_first = p_first.p; ga.p_last = p_last.p; ga.n_children = counters.p; ga.ch_cap = (int)n;
    ga.ch_off = ch_off.p; ga.ch_cnt = ch_cnt.p; ga.ch_pool = pool_cur; ga.pool_ptr = ull.p; ga.err = counters.p + 1;
    ga.stage_rows = stage_rows;
    int n_clustered = 0;
    for (int wv = 1; wv <= n_waves; ++wv) {
      const int64_t cnt = wave_off[wv + 1] - wave_off[wv];
      if (cnt == 0) continue;
      ga.wave_parents = d_wave_parents.p + wave_off[wv];
      // one cluster of CTAs per parent where the wave leaves SMs idle (the largest size that still fits)
      int CS = 1;
      if (cl_mode > 0 || (cl_mode < 0 && n / P >= 32))  // (very short lists: the cluster barriers cost more than they save)
        // (16 CTAs only for long lists: at the fine levels their barriers cost more than the extra CTAs save -
        // N = 8M, D = 2: level 8 of 9 took 423 ms with 16 against 160 ms with 8)
        for (int c = (n / P >= 65536 ? max_cluster : 8); c > 1; c >>= 1)
          if (cnt * c <= ctx->sm_count) { CS = c; break; }
      if (CS > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(cnt * CS));
        cfg.blockDim = dim3(T);
        cfg.dynamicSmemBytes = greedy_smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, greedy_cluster_kernel, ga) == cudaSuccess) {
          ++n_clustered;
        } else {  // e.g. a partitioned device that cannot co-schedule the cluster: one CTA per parent instead
          (void)cudaGetLastError();
          greedy_kernel<<<(unsigned)cnt, T, greedy_smem, st>>>(ga);
        }
      } else {
        greedy_kernel<<<(unsigned)cnt, T, greedy_smem, st>>>(ga);
      }
      CGGP_LAUNCH_CHECK(ctx);
    }
    int host_counters[4] = {0, 0, 0, 0};
    CGGP_CUDA(ctx, cudaMemcpyAsync(host_counters, counters.p, sizeof(host_counters), cudaMemcpyDeviceToHost, st));
    CGGP_CUDA(ctx, cudaStreamSynchronize(st));
    if (trace)
      fprintf(stderr, "covertree level %d: %lld parents, %d waves (%d clustered), %d new nodes, greedy pass %.3f ms\n", level,
              (long long)P, n_waves, n_clustered, host_counters[0], ms_since(t_level));
    const auto t_host = std::chrono::steady_clock::now();
    if (host_counters[1] == 1)
      CGGP_FAIL(ctx, CGGP_ERR_INVALID, "covertree: level %d: a new node took no row of its parent (the reference loops "
                "forever on such input)", level);
    if (host_counters[1] != 0) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "covertree: level %d: node table full", level);
    const int nc = host_counters[0];
    std::vector<double> h_pt((size_t)nc * D);
    std::vector<int> h_next(nc), h_first(P);
    if (nc) {
      CGGP_CUDA(ctx, cudaMemcpyAsync(h_pt.data(), ch_pt.p, sizeof(double) * (size_t)nc * D, cudaMemcpyDeviceToHost, st));
      CGGP_CUDA(ctx, cudaMemcpyAsync(h_next.data(), ch_next.p, sizeof(int) * nc, cudaMemcpyDeviceToHost, st));
    }
    CGGP_CUDA(ctx, cudaMemcpyAsync(h_first.data(), p_first.p, sizeof(int) * P, cudaMemcpyDeviceToHost, st));
    CGGP_CUDA(ctx, cudaStreamSynchronize(st));
    // children in the reference's order: parents in level order, each parent's children in creation order
    std::vector<int> order;
    order.reserve(nc);
    par.child_begin.assign(P, 0);
    par.child_end.assign(P, 0);
    lv.parent.clear();
    for (int64_t i = 0; i < P; ++i) {
      par.child_begin[i] = (int)order.size();
      for (int c = h_first[i]; c >= 0; c = h_next[c]) {
        order.push_back(c);
        lv.parent.push_back((int)i);
      }
      par.child_end[i] = (int)order.size();
    }
    lv.size = nc;
    lv.pts.resize((size_t)nc * D);
    for (int k = 0; k < nc; ++k) memcpy(&lv.pts[(size_t)k * D], &h_pt[(size_t)order[k] * D], sizeof(double) * D);

    // candidate children of every parent = children of its neighbours, in order (covertree.py:103-105, :121-125)
    std::vector<int64_t> cand_off(P + 1, 0);
    for (int64_t i = 0; i < P; ++i) {
      int64_t c = 0;
      for (int q : par.rn[i]) c += par.child_end[q] - par.child_begin[q];
      cand_off[i + 1] = cand_off[i] + c;
    }
    std::vector<int> cand_idx((size_t)cand_off[P]);
    for (int64_t i = 0; i < P; ++i) {
      int64_t w = cand_off[i];
      for (int q : par.rn[i])
        for (int c = par.child_begin[q]; c < par.child_end[q]; ++c) cand_idx[w++] = c;
    }
    // neighbours of the new nodes (covertree.py:106-114)
    lv.rn.assign(nc, {});
    const double reach = nfactor * radius;
    for (int64_t i = 0; i < P; ++i) {
      for (int c = par.child_begin[i]; c < par.child_end[i]; ++c) {
        const double* pc = &lv.pts[(size_t)c * D];
        std::vector<int>& dst = lv.rn[c];
        for (int64_t k = cand_off[i]; k < cand_off[i + 1]; ++k) {
          const int q = cand_idx[k];
          const double* pq = &lv.pts[(size_t)q * D];
          // np.linalg.norm(candidate - child, axis=-1): pairwise sum of the squared differences
          if (std::sqrt(row_sq(pc, pq, D)) <= reach) dst.push_back(q);
        }
      }
    }

    if (trace) fprintf(stderr, "covertree level %d: node order + neighbour lists on the host %.3f ms\n", level, ms_since(t_host));
    const auto t_vor = std::chrono::steady_clock::now();
    if (voronoi) {
      // host copy of the parents' Voronoi extents (their lists lie back to back in vor_cur, in level order)
      std::vector<int64_t> vor_off(P + 1);
      if (level == 1) {
        vor_off[0] = 0;
        vor_off[1] = n;
      } else {
        CGGP_CUDA(ctx, cudaMemcpyAsync(vor_off.data(), seg_off.p, sizeof(int64_t) * P, cudaMemcpyDeviceToHost, st));
        CGGP_CUDA(ctx, cudaStreamSynchronize(st));
        vor_off[P] = n;
      }
      std::vector<int64_t> desc;  // work items: runs of at most VROWS rows of one parent
      for (int64_t i = 0; i < P; ++i)
        for (int64_t s = vor_off[i]; s < vor_off[i + 1]; s += VROWS) {
          desc.push_back(i);
          desc.push_back(s);
          desc.push_back(std::min<int64_t>(VROWS, vor_off[i + 1] - s));
        }
      const int64_t nblk = (int64_t)desc.size() / 3;
      DevBuf<int64_t> d_desc, d_cand_off;
      DevBuf<int> d_cand_idx;
      CGGP_CUDA(ctx, d_desc.alloc(desc.size()));
      CGGP_CUDA(ctx, d_cand_off.alloc(P + 1));
      CGGP_CUDA(ctx, d_cand_idx.alloc(cand_idx.size()));
      CGGP_CUDA(ctx, cudaMemcpyAsync(d_desc.p, desc.data(), sizeof(int64_t) * desc.size(), cudaMemcpyHostToDevice, st));
      CGGP_CUDA(ctx, cudaMemcpyAsync(d_cand_off.p, cand_off.data(), sizeof(int64_t) * (P + 1), cudaMemcpyHostToDevice, st));
      if (!cand_idx.empty())
        CGGP_CUDA(ctx, cudaMemcpyAsync(d_cand_idx.p, cand_idx.data(), sizeof(int) * cand_idx.size(),
                                       cudaMemcpyHostToDevice, st));
      // node points in level order (the greedy pass is over: its table is free)
      if (nc)
        CGGP_CUDA(ctx, cudaMemcpyAsync(ch_pt.p, lv.pts.data(), sizeof(double) * (size_t)nc * D, cudaMemcpyHostToDevice, st));
      int64_t max_c = 0;
      for (int64_t i = 0; i < P; ++i) max_c = std::max(max_c, cand_off[i + 1] - cand_off[i]);
      const int smem_doubles = (int)std::min<int64_t>(max_c * D, vor_smem_doubles);
      if (nblk) {
        voronoi_kernel<<<(unsigned)nblk, VT, sizeof(double) * smem_doubles, st>>>(
            X, ldx, D, vor_cur, d_desc.p, d_cand_off.p, d_cand_idx.p, ch_pt.p, keysA.p, smem_doubles);
        CGGP_LAUNCH_CHECK(ctx);
      }
      int bits = 1;
      while ((1ll << bits) < nc) ++bits;
      size_t tmp_bytes = 0;
      CGGP_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keysA.p, keysB.p, vor_cur, vor_alt, n, 0, bits, st));
      if (tmp_bytes > cub_tmp.n) CGGP_CUDA(ctx, cub_tmp.alloc(tmp_bytes));  // (not expected: sized for 31 key bits)
      CGGP_CUDA(ctx, cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, keysA.p, keysB.p, vor_cur, vor_alt, n, 0, bits, st));
      ctx->launches += 1;
      std::swap(vor_cur, vor_alt);
      bounds_kernel<<<iota_grid, 256, 0, st>>>(keysB.p, n, nc, seg_off.p);
      CGGP_LAUNCH_CHECK(ctx);
      counts_kernel<<<std::max(1, std::min((nc + 255) / 256, 1184)), 256, 0, st>>>(seg_off.p, nc, seg_cnt.p);
      CGGP_LAUNCH_CHECK(ctx);
      // child.data = copy of its Voronoi rows (covertree.py:151-154): the lists the next greedy pass consumes
      CGGP_CUDA(ctx, cudaMemcpyAsync(data_cur, vor_cur, sizeof(int) * n, cudaMemcpyDeviceToDevice, st));
    } else {
      // the rows each new node took are its data (covertree.py:98): lists in the pool, extents by creation id
      std::vector<int64_t> h_off(nc), h_cnt(nc), o2(nc + 1), c2(nc + 1);
      if (nc) {
        CGGP_CUDA(ctx, cudaMemcpyAsync(h_off.data(), ch_off.p, sizeof(int64_t) * nc, cudaMemcpyDeviceToHost, st));
        CGGP_CUDA(ctx, cudaMemcpyAsync(h_cnt.data(), ch_cnt.p, sizeof(int64_t) * nc, cudaMemcpyDeviceToHost, st));
        CGGP_CUDA(ctx, cudaStreamSynchronize(st));
      }
      for (int k = 0; k < nc; ++k) {
        o2[k] = h_off[order[k]];
        c2[k] = h_cnt[order[k]];
      }
      if (nc) {
        CGGP_CUDA(ctx, cudaMemcpyAsync(seg_off.p, o2.data(), sizeof(int64_t) * nc, cudaMemcpyHostToDevice, st));
        CGGP_CUDA(ctx, cudaMemcpyAsync(seg_cnt.p, c2.data(), sizeof(int64_t) * nc, cudaMemcpyHostToDevice, st));
        CGGP_CUDA(ctx, cudaStreamSynchronize(st));
      }
      std::swap(data_cur, pool_cur);
    }
    if (trace) {
      cudaStreamSynchronize(st);
      fprintf(stderr, "covertree level %d: Voronoi pass / list hand-over %.3f ms\n", level, ms_since(t_vor));
    }
  }

  // rows of the last level
  const int64_t m = tree->levels.back().size;
  CGGP_CUDA(ctx, cudaMalloc((void**)&tree->leaf_idx, sizeof(int) * n));
  CGGP_CUDA(ctx, cudaMalloc((void**)&tree->leaf_off, sizeof(int64_t) * std::max<int64_t>(m, 1)));
  CGGP_CUDA(ctx, cudaMalloc((void**)&tree->leaf_cnt, sizeof(int64_t) * std::max<int64_t>(m, 1)));
  CGGP_CUDA(ctx, cudaMemcpyAsync(tree->leaf_idx, data_cur, sizeof(int) * n, cudaMemcpyDeviceToDevice, st));
  if (m) {
    CGGP_CUDA(ctx, cudaMemcpyAsync(tree->leaf_off, seg_off.p, sizeof(int64_t) * m, cudaMemcpyDeviceToDevice, st));
    CGGP_CUDA(ctx, cudaMemcpyAsync(tree->leaf_cnt, seg_cnt.p, sizeof(int64_t) * m, cudaMemcpyDeviceToDevice, st));
  }
  CGGP_CUDA(ctx, cudaStreamSynchronize(st));
  guard.keep = true;
  *out = tree;
  if (trace) fprintf(stderr, "covertree: built in %.3f ms\n", ms_since(t_begin));
  return CGGP_OK;
}

extern "C" int cggp_covertree_destroy(cggp_covertree* tree) {
  delete tree;
  return CGGP_OK;
}
extern "C" int cggp_covertree_num_levels(const cggp_covertree* tree) { return tree ? (int)tree->levels.size() : 0; }
extern "C" int64_t cggp_covertree_level_size(const cggp_covertree* tree, int level) {
  if (!tree || level < 0 || level >= (int)tree->levels.size()) return -1;
  return tree->levels[level].size;
}
extern "C" int cggp_covertree_level_radius(const cggp_covertree* tree, int level, double* host_radius) {
  if (!tree || !host_radius || level < 0 || level >= (int)tree->levels.size()) return CGGP_ERR_INVALID;
  *host_radius = tree->levels[level].radius;
  return CGGP_OK;
}
extern "C" int cggp_covertree_level_points(cggp_ctx* ctx, const cggp_covertree* tree, int level, void* dev_out,
                                           int64_t ldo, int32_t* host_parent) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (!tree || level < 0 || level >= (int)tree->levels.size()) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "covertree: bad level");
  const Level& lv = tree->levels[level];
  if (lv.size && dev_out) {
    if (ldo < tree->D) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "covertree: ldo < D");
    CGGP_CUDA(ctx, cudaMemcpy2DAsync(dev_out, sizeof(double) * ldo, lv.pts.data(), sizeof(double) * tree->D,
                                     sizeof(double) * tree->D, lv.size, cudaMemcpyHostToDevice, ctx->stream));
    CGGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the source is the tree's pageable host copy
  }
  if (host_parent)
    for (int64_t k = 0; k < lv.size; ++k) host_parent[k] = lv.parent[k];
  return CGGP_OK;
}
extern "C" int cggp_covertree_leaf_members(cggp_ctx* ctx, const cggp_covertree* tree, int64_t* dev_offsets,
                                           int64_t* dev_rows) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (!tree || !dev_offsets || !dev_rows) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "covertree: null buffer");
  const int64_t m = tree->levels.back().size;
  exclusive_offsets_kernel<<<1, 32, 0, ctx->stream>>>(tree->leaf_cnt, m, dev_offsets);
  CGGP_LAUNCH_CHECK(ctx);
  if (m) {
    members_kernel<<<(unsigned)m, 128, 0, ctx->stream>>>(tree->leaf_idx, tree->leaf_off, tree->leaf_cnt, m, tree->n,
                                                         dev_offsets, dev_rows);
    CGGP_LAUNCH_CHECK(ctx);
  }
  return CGGP_OK;
}
extern "C" int cggp_covertree_cluster_stats(cggp_ctx* ctx, const cggp_covertree* tree, int dtype, const void* dev_y,
                                            int64_t ldy, void* dev_means, void* dev_counts) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (!tree || !dev_y || !dev_means || !dev_counts || ldy < 1) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "covertree: null buffer");
  const int64_t m = tree->levels.back().size;
  if (m == 0) return CGGP_OK;
  const unsigned grid = (unsigned)((m + 7) / 8);
  if (dtype == CGGP_F64)
    leaf_stats_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double*)dev_y, ldy, tree->leaf_idx, tree->leaf_off,
                                                             tree->leaf_cnt, m, (double*)dev_means, (double*)dev_counts);
  else
    CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "covertree: float64 targets only");
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}
