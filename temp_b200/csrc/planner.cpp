// temp_b200 -- native window planner: the host-side graph builder of the path.
//
// The reference batches DGL graphs per time step inside its forward (dgl.batch, in_degrees, edge_subgraph -- libdgl,
// models/DynamicRGCN.py:66-110) and walks python dictionaries for the history (DynamicRGCN.py:35-54).  Here a whole
// window batch is packed ONCE into the arrays the kernels consume (see temp_b200/planner.py, the readable python
// statement of the same algorithm, which tests/test_host_cpu.py holds this file to, array for array):
//   packed rows step-major, CSR by destination, prev_row / dt maps (the dense history incl. "history forgets"),
//   attention slot rows, the aggregation work lists and the chain partitions of the GRU scan.
// Window construction: models/TKG_Module.py:232-250 (forward), models/BiDynamicRGCN.py:17-49 (backward).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <vector>

#include "temp_b200.h"

namespace {

struct Inst {
  int32_t item, step, dir, time, row0, n, snap;  // dir: 0 'f', 1 'b', 2 'c'
};
struct Seg {
  int32_t kind, step, row0, row1, inst0, inst1;  // kind: 0 hist_f, 1 hist_b, 2 final
};

}  // namespace

struct TempPlan {
  std::vector<int32_t> ent_id, row_time, row_ptr, e_src, e_src_ent, e_rel, prev_a, prev_b, slot_row, scan_parts, agg_rows,
      agg_heavy;
  std::vector<float> norm, dt_a, dt_b;
  std::vector<Inst> insts;
  std::vector<Seg> segs;
  std::vector<int32_t> last_f, last_b;    // per item: instance index of the last history step, -1 = none
  std::vector<int32_t> steps_f, steps_b;  // [L-1][B] instance index per (step, window row j), -1 = none
  TempPlanCounts c;
  bool bi = false, attention = false;
};

namespace {

struct Builder {
  const TempSnapshotView* snaps;
  TempPlan* P;
  int32_t R = 0, E = 0;
  bool bi;

  int32_t add(int32_t si) {
    const TempSnapshotView& s = snaps[si];
    const int32_t row0 = R;
    P->ent_id.insert(P->ent_id.end(), s.node_ids, s.node_ids + s.n_nodes);
    P->row_time.insert(P->row_time.end(), s.n_nodes, s.time);
    P->norm.insert(P->norm.end(), s.norm, s.norm + s.n_nodes);
    {
      const size_t o = P->row_ptr.size();
      P->row_ptr.resize(o + s.n_nodes);
      int32_t* rp = P->row_ptr.data() + o;
      for (int32_t v = 0; v < s.n_nodes; ++v) rp[v] = E + s.row_ptr[v + 1];
      const size_t oe = P->e_src.size();
      P->e_src.resize(oe + s.n_edges);
      P->e_src_ent.resize(oe + s.n_edges);
      int32_t* es = P->e_src.data() + oe;
      int32_t* ee = P->e_src_ent.data() + oe;
      for (int32_t e = 0; e < s.n_edges; ++e) {
        es[e] = s.csr_src[e] + row0;
        ee[e] = s.node_ids[s.csr_src[e]];
      }
    }
    P->e_rel.insert(P->e_rel.end(), s.csr_rel, s.csr_rel + s.n_edges);
    R += s.n_nodes;
    E += s.n_edges;
    return row0;
  }

  // packed row of each node of snapshot `cur` in instance `prev` (same batch item, previous step), -1 if absent;
  // both id lists are sorted: one merge pass
  void match_prev(int32_t cur, int32_t prev_inst, std::vector<int32_t>& out) {
    const TempSnapshotView& c = snaps[cur];
    if (prev_inst < 0) {
      out.insert(out.end(), c.n_nodes, -1);
      return;
    }
    const Inst& pi = P->insts[prev_inst];
    const TempSnapshotView& p = snaps[pi.snap];
    int32_t q = 0;
    for (int32_t v = 0; v < c.n_nodes; ++v) {
      const int32_t id = c.node_ids[v];
      while (q < p.n_nodes && p.node_ids[q] < id) ++q;
      out.push_back(q < p.n_nodes && p.node_ids[q] == id ? pi.row0 + q : -1);
    }
  }
};

void window_rows(const std::vector<int32_t>& order, int32_t n_snaps, int32_t L, bool backward, std::vector<int32_t>& rows) {
  // rows[j * L + k] = snapshot index of window row j at step k, -1 = None padding at the front
  rows.assign(order.size() * L, -1);
  for (size_t j = 0; j < order.size(); ++j) {
    const int32_t p = order[j];
    if (backward) {  // times[p : p + L] reversed, so that the target comes last
      const int32_t len = std::min(L, n_snaps - p);
      for (int32_t i = 0; i < len; ++i) rows[j * L + (L - len) + i] = p + (len - 1 - i);
    } else {         // the last L timestamps <= t
      const int32_t lo = std::max(0, p + 1 - L), len = p + 1 - lo;
      for (int32_t i = 0; i < len; ++i) rows[j * L + (L - len) + i] = lo + i;
    }
  }
}

}  // namespace

extern "C" {

TempPlan* temp_plan_window(const TempSnapshotView* snaps, int32_t n_snaps, const int32_t* targets, int32_t B, int32_t L,
                           int32_t bidirectional, int32_t attention, int32_t scan_tile, int32_t heavy_degree) {
  if (snaps == nullptr || targets == nullptr || n_snaps <= 0 || B <= 0 || L <= 0 || scan_tile == 0) return nullptr;
  for (int32_t i = 0; i < B; ++i)
    if (targets[i] < 0 || targets[i] >= n_snaps) return nullptr;
  TempPlan* P = new TempPlan();
  Builder b{snaps, P};
  b.bi = bidirectional != 0;
  P->row_ptr.push_back(0);

  // snapshots are in graph_dict key order = ascending time for the reference's datasets; sort targets by TIME
  std::vector<int32_t> desc(targets, targets + B), asc;
  std::stable_sort(desc.begin(), desc.end(), [&](int32_t x, int32_t y) { return snaps[x].time > snaps[y].time; });
  asc.assign(desc.rbegin(), desc.rend());
  std::vector<int32_t> fwd, bwd;
  window_rows(desc, n_snaps, L, false, fwd);
  if (b.bi) window_rows(asc, n_snaps, L, true, bwd);

  {  // one allocation per array: rows / edges of all instances are known from the window tables
    size_t tot_r = 0, tot_e = 0;
    auto count = [&](const std::vector<int32_t>& rows, int32_t k0, int32_t k1) {
      for (size_t j = 0; j < static_cast<size_t>(B); ++j)
        for (int32_t k = k0; k < k1; ++k) {
          const int32_t si = rows[j * L + k];
          if (si >= 0) {
            tot_r += snaps[si].n_nodes;
            tot_e += snaps[si].n_edges;
          }
        }
    };
    count(fwd, 0, L);
    if (b.bi) count(bwd, 0, L - 1);
    for (auto* v : {&P->ent_id, &P->row_time, &P->prev_a}) v->reserve(tot_r);
    P->row_ptr.reserve(tot_r + 1);
    P->norm.reserve(tot_r);
    P->dt_a.reserve(tot_r);
    if (b.bi) {
      P->prev_b.reserve(tot_r);
      P->dt_b.reserve(tot_r);
    }
    for (auto* v : {&P->e_src, &P->e_src_ent, &P->e_rel}) v->reserve(tot_e);
    P->agg_rows.reserve(3 * std::min(tot_r, tot_e));
  }

  auto history = [&](const std::vector<int32_t>& rows, int32_t kind, bool flip, std::vector<int32_t>& last,
                     std::vector<int32_t>& steps) {
    last.assign(B, -1);
    steps.assign(static_cast<size_t>(std::max(L - 1, 0)) * B, -1);
    for (int32_t k = 0; k < L - 1; ++k) {
      Seg seg{kind, k, b.R, b.R, static_cast<int32_t>(P->insts.size()), 0};
      std::vector<int32_t> cur(B, -1);
      for (int32_t j = 0; j < B; ++j) {
        const int32_t si = rows[static_cast<size_t>(j) * L + k];
        if (si < 0) continue;
        const int32_t item = flip ? B - 1 - j : j;   // BiDynamicRGCN.py:97-99 (flip)
        b.match_prev(si, last[j], P->prev_a);
        // dt = cur_t - start_time; it only multiplies a non-zero state, i.e. rows whose entity was active at step
        // k-1, where it equals 1
        P->dt_a.insert(P->dt_a.end(), snaps[si].n_nodes, last[j] >= 0 ? 1.0f : static_cast<float>(k));
        if (b.bi) {
          P->prev_b.insert(P->prev_b.end(), snaps[si].n_nodes, -1);
          P->dt_b.insert(P->dt_b.end(), snaps[si].n_nodes, static_cast<float>(k));
        }
        const int32_t row0 = b.add(si);
        cur[j] = static_cast<int32_t>(P->insts.size());
        P->insts.push_back(Inst{item, k, kind, snaps[si].time, row0, snaps[si].n_nodes, si});
      }
      seg.row1 = b.R;
      seg.inst1 = static_cast<int32_t>(P->insts.size());
      for (int32_t j = 0; j < B; ++j) {
        if (cur[j] >= 0) last[j] = cur[j];
        steps[static_cast<size_t>(k) * B + j] = cur[j];
      }
      if (seg.row1 > seg.row0) P->segs.push_back(seg);
    }
  };

  std::vector<int32_t> last_f, last_b;
  history(fwd, 0, false, last_f, P->steps_f);
  if (b.bi) history(bwd, 1, true, last_b, P->steps_b);
  else last_b.assign(B, -1);

  const int32_t n_slots = attention ? (L - 1) * (b.bi ? 2 : 1) : 0;
  Seg seg{2, L - 1, b.R, b.R, static_cast<int32_t>(P->insts.size()), 0};
  for (int32_t i = 0; i < B; ++i) {
    const int32_t si = fwd[static_cast<size_t>(i) * L + (L - 1)];
    b.match_prev(si, last_f[i], P->prev_a);
    P->dt_a.insert(P->dt_a.end(), snaps[si].n_nodes, last_f[i] >= 0 ? 1.0f : static_cast<float>(L - 1));
    if (b.bi) {
      const int32_t lb = last_b[B - 1 - i];
      b.match_prev(si, lb, P->prev_b);
      P->dt_b.insert(P->dt_b.end(), snaps[si].n_nodes, lb >= 0 ? 1.0f : static_cast<float>(L - 1));
    }
    if (attention) {
      const int32_t n = snaps[si].n_nodes;
      const size_t base = P->slot_row.size();
      P->slot_row.resize(base + static_cast<size_t>(n) * n_slots, -1);
      std::vector<int32_t> col;
      for (int32_t s = 0; s < n_slots; ++s) {
        const bool back = s >= L - 1;
        const int32_t k = back ? s - (L - 1) : s;
        const int32_t inst = back ? P->steps_b[static_cast<size_t>(k) * B + (B - 1 - i)] : P->steps_f[static_cast<size_t>(k) * B + i];
        col.clear();
        b.match_prev(si, inst, col);
        for (int32_t v = 0; v < n; ++v) P->slot_row[base + static_cast<size_t>(v) * n_slots + s] = col[v];
      }
    }
    const int32_t row0 = b.add(si);
    P->insts.push_back(Inst{i, L - 1, 2, snaps[si].time, row0, snaps[si].n_nodes, si});
  }
  seg.row1 = b.R;
  seg.inst1 = static_cast<int32_t>(P->insts.size());
  P->segs.push_back(seg);
  P->last_f = last_f;
  P->last_b.assign(B, -1);
  if (b.bi)
    for (int32_t i = 0; i < B; ++i) P->last_b[i] = last_b[B - 1 - i];

  // ---- aggregation work lists: (row, first edge, end edge) of the rows with in-edges, split at heavy_degree --------
  // (heavy_degree <= 0: four times the mean in-degree of the rows that have in-edges, within [8, 64] --
  // temp_b200/planner.py::agg_heavy_degree)
  if (heavy_degree <= 0) {
    int64_t nz = 0;
    for (int32_t r = 0; r < b.R; ++r) nz += P->row_ptr[r + 1] > P->row_ptr[r];
    const int64_t e = b.R > 0 ? P->row_ptr[b.R] : 0;
    heavy_degree = nz > 0 ? static_cast<int32_t>(std::min<int64_t>(64, std::max<int64_t>(8, (4 * e) / nz))) : 8;
  }
  for (int32_t r = 0; r < b.R; ++r) {
    const int32_t p0 = P->row_ptr[r], p1 = P->row_ptr[r + 1];
    if (p1 <= p0) continue;
    std::vector<int32_t>& dst = (p1 - p0 > heavy_degree) ? P->agg_heavy : P->agg_rows;
    dst.push_back(r);
    dst.push_back(p0);
    dst.push_back(p1);
  }

  // ---- chain partitions (temp_b200/planner.py::chain_partitions) ----------------------------------------------------
  const int32_t n_seg = static_cast<int32_t>(P->segs.size());
  int32_t n_parts = 0;
  // scan_tile < 0: |scan_tile| rows per partition step, widened to 64 when that leaves more partitions than two rounds of the
  // scan kernel's pipelines (2 x 33 clusters x 2): the scan is then throughput bound and 25 % fewer, fuller tile steps win
  // (x16 bench shape: 367 -> 334 us), while in the latency regime the narrower tile is faster (x1: 45 against 51 us)
  const bool auto_tile = scan_tile < 0;
  if (auto_tile) scan_tile = -scan_tile;
  int32_t used_tile = scan_tile;
  if (!attention) {
    struct Part {
      std::vector<int32_t> rg;
      int64_t rows;
    };
    std::vector<Part> parts;
    std::vector<std::vector<std::pair<int32_t, int32_t>>> per_item(B);   // (segment index, instance index)
    for (int32_t g = 0; g < n_seg; ++g)
      for (int32_t ii = P->segs[g].inst0; ii < P->segs[g].inst1; ++ii) per_item[P->insts[ii].item].push_back({g, ii});
    const int64_t big = std::numeric_limits<int64_t>::max();
    auto cut_parts = [&](int32_t tile) {
      parts.clear();
      for (int32_t item = 0; item < B; ++item) {
        const auto& li = per_item[item];
        const size_t K = li.size();
        std::vector<int32_t> pos(K, 0);
        while (true) {
          bool any = false;
          for (size_t k = 0; k < K; ++k) any = any || pos[k] < P->insts[li[k].second].n;
          if (!any) break;
          // the first entity id that would make some instance exceed `tile` rows; everything below it joins
          int64_t cut = big;
          for (size_t k = 0; k < K; ++k) {
            const Inst& in = P->insts[li[k].second];
            if (pos[k] + tile < in.n) cut = std::min<int64_t>(cut, snaps[in.snap].node_ids[pos[k] + tile]);
          }
          Part pt;
          pt.rg.assign(static_cast<size_t>(n_seg) * 2, 0);
          pt.rows = 0;
          for (size_t k = 0; k < K; ++k) {
            const Inst& in = P->insts[li[k].second];
            const int32_t* ids = snaps[in.snap].node_ids;
            const int32_t nw = cut == big ? in.n
                                          : static_cast<int32_t>(std::lower_bound(ids, ids + in.n, static_cast<int32_t>(cut)) - ids);
            pt.rg[2 * li[k].first] = in.row0 + pos[k];
            pt.rg[2 * li[k].first + 1] = in.row0 + nw;
            pt.rows += nw - pos[k];
            pos[k] = nw;
          }
          parts.push_back(std::move(pt));
        }
      }
    };
    cut_parts(scan_tile);
    if (auto_tile && scan_tile < 64 && parts.size() > 132) {
      used_tile = 64;
      cut_parts(used_tile);
    }
    std::stable_sort(parts.begin(), parts.end(), [](const Part& x, const Part& y) { return x.rows > y.rows; });
    for (const Part& pt : parts) P->scan_parts.insert(P->scan_parts.end(), pt.rg.begin(), pt.rg.end());
    n_parts = static_cast<int32_t>(parts.size());
  }

  P->bi = b.bi;
  P->attention = attention != 0;
  TempPlanCounts& c = P->c;
  memset(&c, 0, sizeof(c));
  c.rows = b.R;
  c.edges = b.E;
  c.n_segments = n_seg;
  c.n_instances = static_cast<int32_t>(P->insts.size());
  c.n_parts = n_parts;
  c.n_agg_rows = static_cast<int32_t>(P->agg_rows.size() / 3);
  c.n_agg_heavy = static_cast<int32_t>(P->agg_heavy.size() / 3);
  c.n_slots = n_slots;
  c.batch = B;
  c.seq_len = L;
  c.scan_tile = used_tile;
  return P;
}

void temp_plan_destroy(TempPlan* plan) { delete plan; }

int temp_plan_counts(const TempPlan* plan, TempPlanCounts* out) {
  if (plan == nullptr || out == nullptr) return TEMP_EINVAL;
  *out = plan->c;
  return TEMP_OK;
}

const void* temp_plan_array(const TempPlan* plan, int32_t which, int64_t* n_bytes) {
  if (plan == nullptr || n_bytes == nullptr) return nullptr;
  const void* p = nullptr;
  int64_t n = 0;
#define TEMP_ARR(ID, V)                                  \
  case ID:                                               \
    p = plan->V.data();                                  \
    n = static_cast<int64_t>(plan->V.size() * sizeof(plan->V[0])); \
    break;
  switch (which) {
    TEMP_ARR(TEMP_PLAN_ENT_ID, ent_id)
    TEMP_ARR(TEMP_PLAN_ROW_TIME, row_time)
    TEMP_ARR(TEMP_PLAN_NORM, norm)
    TEMP_ARR(TEMP_PLAN_ROW_PTR, row_ptr)
    TEMP_ARR(TEMP_PLAN_E_SRC, e_src)
    TEMP_ARR(TEMP_PLAN_E_SRC_ENT, e_src_ent)
    TEMP_ARR(TEMP_PLAN_E_REL, e_rel)
    TEMP_ARR(TEMP_PLAN_PREV_A, prev_a)
    TEMP_ARR(TEMP_PLAN_DT_A, dt_a)
    TEMP_ARR(TEMP_PLAN_PREV_B, prev_b)
    TEMP_ARR(TEMP_PLAN_DT_B, dt_b)
    TEMP_ARR(TEMP_PLAN_SLOT_ROW, slot_row)
    TEMP_ARR(TEMP_PLAN_SCAN_PARTS, scan_parts)
    TEMP_ARR(TEMP_PLAN_AGG_ROWS, agg_rows)
    TEMP_ARR(TEMP_PLAN_AGG_HEAVY, agg_heavy)
    TEMP_ARR(TEMP_PLAN_INSTANCES, insts)
    TEMP_ARR(TEMP_PLAN_SEGMENTS, segs)
    TEMP_ARR(TEMP_PLAN_LAST_F, last_f)
    TEMP_ARR(TEMP_PLAN_LAST_B, last_b)
    TEMP_ARR(TEMP_PLAN_STEPS_F, steps_f)
    TEMP_ARR(TEMP_PLAN_STEPS_B, steps_b)
    default:
      return nullptr;
  }
#undef TEMP_ARR
  *n_bytes = n;
  return p;
}

// The device blob of a plan: arrays TEMP_PLAN_ENT_ID .. TEMP_PLAN_AGG_HEAVY back to back, each start aligned to `align`
// bytes (the layout temp_b200/planner.py::WindowPlan.blob_layout produces); offsets[i] = -1 for arrays the plan does not
// have (prev_b / dt_b of uni-directional plans, slot_row without attention, scan_parts with it).
int64_t temp_plan_blob_layout(const TempPlan* plan, int64_t* offsets, int64_t* sizes, int32_t align) {
  if (plan == nullptr || offsets == nullptr || sizes == nullptr || align <= 0) return -1;
  int64_t off = 0;
  for (int32_t i = 0; i <= TEMP_PLAN_AGG_HEAVY; ++i) {
    const bool absent = ((i == TEMP_PLAN_PREV_B || i == TEMP_PLAN_DT_B) && !plan->bi) ||
                        (i == TEMP_PLAN_SLOT_ROW && !plan->attention) || (i == TEMP_PLAN_SCAN_PARTS && plan->attention);
    int64_t nb = 0;
    if (absent || temp_plan_array(plan, i, &nb) == nullptr && nb != 0) {
      offsets[i] = -1;
      sizes[i] = 0;
      continue;
    }
    temp_plan_array(plan, i, &nb);
    offsets[i] = off;
    sizes[i] = nb;
    off = (off + nb + align - 1) / align * align;
  }
  return off;
}

int temp_plan_write_blob(const TempPlan* plan, uint8_t* dst, int32_t align) {
  if (plan == nullptr || dst == nullptr) return TEMP_EINVAL;
  int64_t offsets[TEMP_PLAN_AGG_HEAVY + 1], sizes[TEMP_PLAN_AGG_HEAVY + 1];
  if (temp_plan_blob_layout(plan, offsets, sizes, align) < 0) return TEMP_EINVAL;
  for (int32_t i = 0; i <= TEMP_PLAN_AGG_HEAVY; ++i) {
    if (offsets[i] < 0 || sizes[i] == 0) continue;
    int64_t nb = 0;
    const void* src = temp_plan_array(plan, i, &nb);
    memcpy(dst + offsets[i], src, static_cast<size_t>(nb));
  }
  return TEMP_OK;
}

// ---- filtered negative sampling on the reference's own random stream --------------------------------------------------
// The reference draws `np.random.randint(num_entities, size=negative_rate)` once per round (utils/CorrptTriples.py:61-85)
// from NumPy's global legacy generator.  That stream is frozen by NumPy's compatibility policy: MT19937, and for a range
// below 2^32 one tempered 32-bit output per attempt, masked to the next power of two and redrawn while above the range
// (numpy/random/src/distributions/distributions.c, buffered_bounded_masked_uint32 under RandomState's use_masked mode).
// temp_negative_sample continues that generator from the state np.random.get_state() reports and hands the advanced
// state back, so the candidates AND every later draw of the process are bit-identical to the reference's.
//
// Per slot it restates the reference's loop (rounds filtered by np.in1d(cand, true_ids, assume_unique=True, invert=True),
// concatenated, truncated) INCLUDING which of numpy's three in1d algorithms the call takes
// (numpy/lib/_arraysetops_impl.py, numpy 2.x): the integer table and the per-value loop are plain membership tests; the
// merge-sort path with assume_unique=True additionally drops every candidate that has an equal candidate LATER in the
// same round (the stable sort places it before its twin, and the "next element differs" test then fails).
namespace {
struct Mt19937 {
  uint32_t* key;
  int32_t pos;
  void refill() {
    constexpr int N = 624, M = 397;
    constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;
    int kk = 0;
    uint32_t y;
    for (; kk < N - M; ++kk) {
      y = (key[kk] & kUpper) | (key[kk + 1] & kLower);
      key[kk] = key[kk + M] ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
    }
    for (; kk < N - 1; ++kk) {
      y = (key[kk] & kUpper) | (key[kk + 1] & kLower);
      key[kk] = key[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
    }
    y = (key[N - 1] & kUpper) | (key[0] & kLower);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
    pos = 0;
  }
  uint32_t next() {
    if (pos >= 624) refill();
    uint32_t y = key[pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};
}  // namespace

int64_t temp_negative_sample(uint32_t* mt_key, int32_t* mt_pos, int64_t num_entities, int32_t neg, const int64_t* fptr,
                             const int64_t* fids, int32_t n_slots, int32_t sort_min_f, int64_t* out_even, int64_t* out_odd,
                             int64_t out_stride) {
  if (mt_key == nullptr || mt_pos == nullptr || fptr == nullptr || (fids == nullptr && fptr[n_slots] > 0) ||
      out_even == nullptr || out_odd == nullptr || neg <= 0 || n_slots < 0 || num_entities <= 0 ||
      num_entities > 0xFFFFFFFFll || *mt_pos < 0 || *mt_pos > 624)
    return TEMP_EINVAL;
  Mt19937 mt{mt_key, *mt_pos};
  const uint32_t rng = static_cast<uint32_t>(num_entities - 1);
  uint32_t mask = rng;
  mask |= mask >> 1;
  mask |= mask >> 2;
  mask |= mask >> 4;
  mask |= mask >> 8;
  mask |= mask >> 16;
  int64_t rounds = 0;
  std::vector<int64_t> forb, cand(static_cast<size_t>(neg));
  std::vector<uint8_t> keep(static_cast<size_t>(neg));
  std::vector<int32_t> order(static_cast<size_t>(neg));
  for (int32_t s = 0; s < n_slots; ++s) {
    const int64_t f = fptr[s + 1] - fptr[s];
    forb.assign(fids + fptr[s], fids + fptr[s + 1]);
    std::sort(forb.begin(), forb.end());
    bool sort_path = false;
    if (f > 0) {
      const int64_t range = forb.back() - forb.front();
      const bool table = range <= 6 * (static_cast<int64_t>(neg) + f);
      sort_path = !table && f >= sort_min_f;
    }
    int64_t* dst = ((s & 1) ? out_odd : out_even) + static_cast<int64_t>(s >> 1) * out_stride;
    int32_t filled = 0;
    const int64_t* fb = forb.data();
    while (filled < neg) {
      ++rounds;
      if (!sort_path) {
        // one round of np.random.randint(num_entities, size=neg): every draw is taken from the stream, the ones outside
        // the filter list are kept while the slot still has room
        for (int32_t i = 0; i < neg; ++i) {
          uint32_t v = 0;
          if (rng != 0) {
            do {
              v = mt.next() & mask;
            } while (v > rng);
          }
          const int64_t c = static_cast<int64_t>(v);
          bool in = false;
          if (f <= 8) {
            for (int64_t k = 0; k < f; ++k) in |= fb[k] == c;
          } else {
            in = std::binary_search(forb.begin(), forb.end(), c);
          }
          if (!in && filled < neg) dst[filled++] = c;
        }
        continue;
      }
      for (int32_t i = 0; i < neg; ++i) {
        uint32_t v = 0;
        if (rng != 0) {
          do {
            v = mt.next() & mask;
          } while (v > rng);
        }
        cand[static_cast<size_t>(i)] = static_cast<int64_t>(v);
        keep[static_cast<size_t>(i)] = std::binary_search(forb.begin(), forb.end(), static_cast<int64_t>(v)) ? 0 : 1;
      }
      // drop every candidate with an equal candidate later in the round
      for (int32_t i = 0; i < neg; ++i) order[static_cast<size_t>(i)] = i;
      std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cand[static_cast<size_t>(a)] < cand[static_cast<size_t>(b)]; });
      for (int32_t k = 0; k + 1 < neg; ++k)
        if (cand[static_cast<size_t>(order[static_cast<size_t>(k)])] == cand[static_cast<size_t>(order[static_cast<size_t>(k + 1)])])
          keep[static_cast<size_t>(order[static_cast<size_t>(k)])] = 0;
      for (int32_t i = 0; i < neg && filled < neg; ++i)
        if (keep[static_cast<size_t>(i)]) dst[filled++] = cand[static_cast<size_t>(i)];
    }
  }
  *mt_pos = mt.pos;
  return rounds;
}

}  // extern "C"
