// qe_kernels.cuh -- CUDA kernels of the engine (sm_100a): select, exact sequential TD update, env steps, fused loop.
//
// Table layout in HBM.  Q is dense: one row of 8*LPR floats per state (A=16: 64 B, so the 1M-state table of config 3
// is 64 MB and lives in L2).  The writer lists of this file's exact update live in a SEPARATE array (`info`, three
// times the row size per state, allocated the first time a writer-list kernel runs): round 1 kept them inside a
// 256-byte row block, which made every row gather of every kernel stride over metadata only this form reads.
//
// Exact sequential TD update in parallel (DESIGN.md "TD update"):
//   The reference applies agents 0..N-1 one after the other (QLO:806-817), so agent i must see every write of
//   agents j<i -- to its own cell Q[s_i,a_i] AND to the row Q[s'_i,:] it bootstraps from.
//   Phase 1 (insert)  : every agent captures p_i = Q0[s_i,a_i] and registers itself as a writer of row s_i:
//                       count = atomicAdd on the row's epoch-stamped counter, entry {agent, action} stored inline
//                       in the row block (overflow: linked list through node[]).
//   -- grid-wide barrier --
//   Phase 2 (resolve) : agent i scans the writer entries of s_i and s'_i.  For every cell it needs it finds the
//                       latest writer j<i; the value "just before i" is slot[j] if such a writer exists, a writer's
//                       captured p (== Q0; the table may already hold a later commit) if the cell is written only
//                       by agents >= i, and the table itself if nobody writes the cell this step.
//                       v_i = p + lr*((r + gamma*m) - p) is published in slot[i] = {epoch, v_i} (one 64-bit store);
//                       waiting agents poll their predecessors' slots.  The dependency graph is a DAG in agent
//                       order and every warp processes its agents in increasing order, so the smallest unresolved
//                       agent can always proceed (cooperative launch => all CTAs are co-resident).
//                       The last writer of a cell commits v to the table.
//   Everything is stamped with a 32-bit epoch (global step counter), so nothing is cleared between steps.
//
// Thread mapping: scalar work (hashing, env step, writer scans, agent-array traffic) is one lane per agent, 32
// consecutive agents per warp iteration (coalesced); row gathers are transposed so that LPA lanes fetch one row with
// one 16-byte load each (one 128-byte line per row, not one per lane).
#pragma once
#include <cooperative_groups.h>

#include "qe_common.cuh"

namespace qe {
namespace cg = cooperative_groups;

constexpr uint32_t kNone = 0xFFFFFFu;       // list terminator (24-bit agent index)
constexpr uint32_t kSpillCap = 4096;        // entries per spill array
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr int kErrInvalidMove = 1, kErrEmpty = 2, kErrTimeout = 4;
constexpr uint64_t kTimeoutNs = 4000000000ull;  // a TD update that has not drained after 4 s is reported, not waited for

struct Table {
    float* q;          // [S][ld] dense Q rows
    int ld;            // floats per row (8 * LPR: whole 32-byte sectors)
    int A;             // actions
    uint32_t* info;    // [S][info_ld] writer info of the writer-list form (nullptr until first used)
    int info_ld;       // words per state in info[]
    int inline_cap;    // writer entries stored inline in a state's info block (>= 4); the two words after them hold the spill descriptor
    uint32_t* spill;   // [spill_slots][kSpillCap] contiguous writer entries of crowded rows (beyond the inline ones)
    int spill_slots;
    int* spill_next;   // [2] slots handed out in this step (by epoch parity; the other one is reset meanwhile)
    uint32_t* node;    // [cap] overflow list: (action << 24) | next24
    uint64_t* slot;    // [cap] (epoch << 32) | float bits of v_i
    float* tr_p;       // [cap] Q0[s_i, a_i] captured before any commit of this step
    uint32_t* rec;     // [cap][8] deferred records (second pass of the TD update)
    uint32_t* dmask;   // [cap/32] per tile of 32 agents: which agents were deferred (record holds all they wait for)
    uint32_t* smask;   // [cap/32] ... and which need the in-order pass (crowded rows)
    int* err;          // device error flags
    int hold;          // 1: the TD update publishes values but leaves the table untouched (sharded mode iterates)
    uint8_t* later_buf; // [cap] hold mode: 1 = a later agent writes the same cell (this one must not commit)
};
// writer info words: [0] count, [1] epoch (one u64, atomics), [2] overflow head idx, [3] its epoch (one u64),
//                    [4 .. 4+inline_cap) entries (agent24 | action << 24)
__device__ __forceinline__ uint32_t* row_info(const Table& T, int s) {
    return T.info + (size_t)s * T.info_ld;
}

__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ------------------------------------------------------------------ phase 1: register agent i as a writer of row s
// Entry `ticket` of a row's writer list.  The first inline_cap entries live in the row block.  The agent that draws
// ticket == inline_cap claims a spill array for the row and publishes {slot, epoch} in the descriptor behind the inline
// entries; later arrivals wait for the descriptor (its writer is resident and waits for nobody) and store at
// spill[slot][ticket - inline_cap] -- crowded rows stay contiguous and are read with whole-sector loads.  Only what
// does not fit there (no slot left, or more than kSpillCap entries) goes to the per-agent linked list.
__device__ __forceinline__ void list_insert(const Table& T, unsigned long long* info, int i, int a, uint32_t epoch) {
    const unsigned long long old = atomicExch(info + 1, ((unsigned long long)epoch << 32) | (unsigned long long)i);
    T.node[i] = ((uint32_t)a << 24) | (((uint32_t)(old >> 32) == epoch) ? ((uint32_t)old & kNone) : kNone);
}
__device__ __forceinline__ void store_entry(const Table& T, unsigned long long* info, uint32_t ticket, int i, int a, uint32_t epoch) {
    uint32_t* words = reinterpret_cast<uint32_t*>(info);
    const uint32_t entry = (uint32_t)i | ((uint32_t)a << 24);
    const uint32_t cap = (uint32_t)T.inline_cap;
    if (ticket < cap) {
        words[4 + ticket] = entry;
        return;
    }
    uint64_t* desc = reinterpret_cast<uint64_t*>(words + 4 + cap);
    uint32_t slot = kNoSlot;
    if (ticket == cap) {  // the claiming lane publishes before any lane of its own warp starts to wait (next statement)
        const int got = atomicAdd(T.spill_next + (epoch & 1u), 1);
        slot = got < T.spill_slots ? (uint32_t)got : kNoSlot;
        st_relaxed_u64(desc, ((uint64_t)epoch << 32) | slot);
    }
    if (ticket > cap) {
        uint64_t d;
        do { d = ld_relaxed_u64(desc); } while ((uint32_t)(d >> 32) != epoch);
        slot = (uint32_t)d;
    }
    const uint32_t idx = ticket - cap;
    if (slot != kNoSlot && idx < kSpillCap) T.spill[(size_t)slot * kSpillCap + idx] = entry;
    else list_insert(T, info, i, a, epoch);
}
__device__ __forceinline__ void row_insert(const Table& T, int i, int s, int a, float p, uint32_t epoch) {
    T.tr_p[i] = p;
    unsigned long long* info = reinterpret_cast<unsigned long long*>(row_info(T, s));
    atomicMax(info, (unsigned long long)epoch << 32);  // a stale (older-epoch) counter restarts at {epoch, 0}
    const uint32_t c = (uint32_t)atomicAdd(info, 1ull);
    store_entry(T, info, c, i, a, epoch);
}

struct RowWriters {  // snapshot of a row's writer info (taken after the grid barrier)
    const uint32_t* iw;
    uint4 e;          // first four inline entries
    uint32_t count;   // writers of this row in the current step (0 if the info is stale)
    uint32_t ovf;     // overflow list head or kNone
};
__device__ __forceinline__ RowWriters load_writers(const Table& T, int s, uint32_t epoch) {
    RowWriters w;
    w.iw = row_info(T, s);
    const uint4 h = __ldcg(reinterpret_cast<const uint4*>(w.iw));
    w.count = (h.y == epoch) ? h.x : 0u;
    w.ovf = (h.w == epoch) ? (h.z & kNone) : kNone;
    w.e = make_uint4(0, 0, 0, 0);
    if (w.count) w.e = __ldcg(reinterpret_cast<const uint4*>(w.iw) + 1);
    return w;
}
// writers beyond the first four: the rest of the inline entries (whole 32-byte sectors of eight), then the overflow list
template <typename F>
__device__ __forceinline__ void for_each_writer_from4(const Table& T, const RowWriters& w, F f) {
    if (w.count > 4) {
        const uint32_t ninl = min(w.count, (uint32_t)T.inline_cap);
        for (uint32_t base = 4; base < ninl; base += 8) {
            uint32_t e[8];
            asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]), "=r"(e[4]), "=r"(e[5]), "=r"(e[6]), "=r"(e[7])
                         : "l"(w.iw + 4 + base));
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (base + t < ninl) f(e[t]);
        }
        if (w.count > (uint32_t)T.inline_cap) {
            // spill array (whole sectors, two in flight), then whatever had to go to the linked list
            const uint64_t d = ld_relaxed_u64(reinterpret_cast<const uint64_t*>(w.iw + 4 + T.inline_cap));
            const uint32_t slot = (uint32_t)d;
            uint32_t nsp = 0;
            if (slot != kNoSlot) {
                nsp = min(w.count - (uint32_t)T.inline_cap, kSpillCap);
                const uint32_t* sp = T.spill + (size_t)slot * kSpillCap;
                for (uint32_t base = 0; base < nsp; base += 16) {
                    uint32_t e[16];
                    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]), "=r"(e[4]), "=r"(e[5]), "=r"(e[6]), "=r"(e[7])
                                 : "l"(sp + base));
                    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(e[8]), "=r"(e[9]), "=r"(e[10]), "=r"(e[11]), "=r"(e[12]), "=r"(e[13]), "=r"(e[14]), "=r"(e[15])
                                 : "l"(sp + base + 8));
#pragma unroll
                    for (int t = 0; t < 16; ++t)
                        if (base + t < nsp) f(e[t]);
                }
            }
            if (w.count - (uint32_t)T.inline_cap > nsp) {
                uint32_t j = w.ovf;
                while (j != kNone) {
                    const uint32_t nd = __ldcg(T.node + j);
                    f(j | (nd & 0xFF000000u));
                    j = nd & kNone;
                }
            }
        }
    }
}
template <typename F>
__device__ __forceinline__ void for_each_writer(const Table& T, const RowWriters& w, F f) {
    if (w.count > 0) f(w.e.x);
    if (w.count > 1) f(w.e.y);
    if (w.count > 2) f(w.e.z);
    if (w.count > 3) f(w.e.w);
    for_each_writer_from4(T, w, f);
}

// ------------------------------------------------------------------ 256-bit row / writer-info loads
// sm_100a has 256-bit global loads (SASS LDG.E.ENL2.256): one lane moves one 32-byte sector.  A Q row is
// LPR = ceil(A/8) sectors (1, 2 or 4 lanes), the first sector of the writer info holds {count, epoch, overflow
// head, its epoch, e0..e3}.
struct __align__(32) F8 { float v[8]; };
struct __align__(32) U8 { uint32_t w[8]; };
__device__ __forceinline__ F8 ld_row8(const float* p) {
    F8 r;
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ U8 ld_info8(const uint32_t* p) {
    U8 r;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_u8(uint32_t* p, const U8& r) {
    asm volatile("st.global.cg.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]), "r"(r.w[3]),
                 "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7])
                 : "memory");
}
// k-th (0-based) set bit of m, k < popc(m)
__device__ __forceinline__ int kth_set32(uint32_t m, int k) {
    int pos = 0, c;
    c = __popc(m & 0xFFFFu); if (k >= c) { k -= c; pos += 16; m >>= 16; }
    c = __popc(m & 0xFFu);   if (k >= c) { k -= c; pos += 8;  m >>= 8; }
    c = __popc(m & 0xFu);    if (k >= c) { k -= c; pos += 4;  m >>= 4; }
    c = __popc(m & 0x3u);    if (k >= c) { k -= c; pos += 2;  m >>= 2; }
    c = (int)(m & 1u);       if (k >= c) { pos += 1; }
    return pos;
}
__device__ __forceinline__ float max8(const F8& v, uint32_t legal) {
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x = ((legal >> k) & 1u) ? v.v[k] : -INFINITY;
        m = fmax_plain(m, x);
    }
    return m;
}
__device__ __forceinline__ uint32_t tie8(const F8& v, uint32_t legal, float m) {
    uint32_t t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t |= (v.v[k] == m) ? (1u << k) : 0u;
    return t & legal;
}
__device__ __forceinline__ float sel8(const F8& v, int k) {  // v.v[k & 7] without local memory
    const float a = (k & 1) ? v.v[1] : v.v[0], b = (k & 1) ? v.v[3] : v.v[2], c = (k & 1) ? v.v[5] : v.v[4], d = (k & 1) ? v.v[7] : v.v[6];
    const float e = (k & 2) ? b : a, f = (k & 2) ? d : c;
    return (k & 4) ? f : e;
}

// Transposed row gathers: every lane owns one agent; in round q the LPR lanes of group g serve the agent owned by
// lane q*(32/LPR)+g, one 32-byte sector each (one L1 wavefront per row).  All 32 lanes must call.
template <int LPR>
struct RowGather {
    F8 v[LPR];
    // issue the loads of all rounds (rows whose owner passes want=false are skipped)
    __device__ __forceinline__ void issue(const Table& T, int s, bool want) {
        constexpr int G = 32 / LPR;
        const int lane = threadIdx.x & 31, l = lane & (LPR - 1), g = lane / LPR;
#pragma unroll
        for (int q = 0; q < LPR; ++q) {
            const int src = q * G + g;
            const int ss = (LPR == 1) ? s : __shfl_sync(kFull, s, src);
            const bool ww = (LPR == 1) ? want : (__shfl_sync(kFull, (int)want, src) != 0);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[q].v[k] = 0.0f;
            if (ww) v[q] = ld_row8(T.q + (size_t)ss * T.ld + 8 * l);
        }
    }
    // masked max of the owner's row over `legal`
    __device__ __forceinline__ float row_max(uint32_t legal) const {
        if (LPR == 1) return max8(v[0], legal & 0xFFu);
        constexpr int G = 32 / LPR;
        const int lane = threadIdx.x & 31, l = lane & (LPR - 1), g = lane / LPR;
        float res = -INFINITY;
#pragma unroll
        for (int q = 0; q < LPR; ++q) {
            const uint32_t my = (__shfl_sync(kFull, legal, q * G + g) >> (8 * l)) & 0xFFu;
            float m = max8(v[q], my);
#pragma unroll
            for (int d = LPR / 2; d > 0; d >>= 1) m = fmax_plain(m, __shfl_xor_sync(kFull, m, d));
            const float back = __shfl_sync(kFull, m, (lane & (G - 1)) * LPR);
            if (lane / G == q) res = back;
        }
        return res;
    }
    // masked max and the set of legal actions attaining it (exact == on fp32, QLO:346 / :469)
    __device__ __forceinline__ void row_max_tie(uint32_t legal, float& m_out, uint32_t& tie_out) const {
        if (LPR == 1) {
            m_out = max8(v[0], legal & 0xFFu);
            tie_out = tie8(v[0], legal & 0xFFu, m_out);
            return;
        }
        constexpr int G = 32 / LPR;
        const int lane = threadIdx.x & 31, l = lane & (LPR - 1), g = lane / LPR;
        m_out = -INFINITY;
        tie_out = 0u;
#pragma unroll
        for (int q = 0; q < LPR; ++q) {
            const uint32_t my = (__shfl_sync(kFull, legal, q * G + g) >> (8 * l)) & 0xFFu;
            float m = max8(v[q], my);
#pragma unroll
            for (int d = LPR / 2; d > 0; d >>= 1) m = fmax_plain(m, __shfl_xor_sync(kFull, m, d));
            uint32_t t = tie8(v[q], my, m) << (8 * l);
#pragma unroll
            for (int d = LPR / 2; d > 0; d >>= 1) t |= __shfl_xor_sync(kFull, t, d);
            const float mb = __shfl_sync(kFull, m, (lane & (G - 1)) * LPR);
            const uint32_t tb = __shfl_sync(kFull, t, (lane & (G - 1)) * LPR);
            if (lane / G == q) { m_out = mb; tie_out = tb; }
        }
    }
    // Q[s, a] of the owner's row for the owner's action a (a < 0 -> unspecified)
    __device__ __forceinline__ float value_of(int a) const {
        if (LPR == 1) return sel8(v[0], a);
        constexpr int G = 32 / LPR;
        const int lane = threadIdx.x & 31, g = lane / LPR;
        float res = 0.0f;
#pragma unroll
        for (int q = 0; q < LPR; ++q) {
            const int aa = __shfl_sync(kFull, a, q * G + g);
            const float x = sel8(v[q], aa);
            // the sector holding action aa sits in lane (aa >> 3) of the serving group
            const float back = __shfl_sync(kFull, x, (lane & (G - 1)) * LPR + ((a >> 3) & (LPR - 1)));
            if (lane / G == q) res = back;
        }
        return res;
    }
};

// masked epsilon-greedy pick from the row statistics (QLO:304-348, 432-470)
__device__ __forceinline__ int pick_action(int num_actions, uint32_t valid, uint32_t tie, bool explore, bool empty_all,
                                           uint32_t bits_pick) {
    uint32_t cand = explore ? valid : tie;
    if (!explore && valid == 0u && empty_all)  // QLO:467-470: every masked value is -inf -> all actions tie
        cand = num_actions >= 32 ? 0xFFFFFFFFu : ((1u << num_actions) - 1u);
    const int cnt = __popc(cand);
    return cnt > 0 ? kth_set32(cand, (int)pick(bits_pick, (uint32_t)cnt)) : -1;  // choice(cand), ascending action order
}

// ------------------------------------------------------------------ phase 2: scan, then resolve + commit
// What agent i needs from the other agents of this step, found by scanning the writer lists of s_i and s'_i:
//   pj      latest earlier writer of the same cell (s_i, a_i), or -1
//   later   some later agent writes the same cell (then i does not commit)
//   m       max over the legal cells of s'_i whose value "just before i" is already known: cells nobody writes
//           (table) and cells written only by agents >= i (Q0 as captured by one of them in tr_p)
//   d0..d3  earlier writers of the remaining legal cells (latest one per cell) whose published value is needed, or -1
//   dyn     crowded bootstrap row (> 4 writers): actions with an earlier writer; the writers sit in the shared-memory table best[]
struct Scan {
    float m;
    int pj, d0, d1, d2, d3;
    uint32_t later, dyn, slow;
};

// One lane per agent; all 32 lanes must call (transposed gathers inside).  Rows with at most four writers are handled
// branch-free from the first info sector; crowded rows walk the whole list (`best`: this thread's column of a
// shared-memory table [8*LPR actions][256 threads]).  SLOW_OK = false: the earlier writers must fit d0..d3, else the
// agent is flagged Scan::slow and left to the in-order pass, which keeps best[] alive while it polls (SLOW_OK = true).
template <int LPR, bool SLOW_OK>
__device__ __forceinline__ Scan scan_writers(const Table& T, int* best, bool active, int i, int s, int a, int s2, bool term,
                                             uint32_t mask2, uint32_t epoch) {
    const bool boot = active && !term;
    U8 I1, I2;
#pragma unroll
    for (int k = 0; k < 8; ++k) I1.w[k] = I2.w[k] = 0u;
    if (active) I1 = ld_info8(row_info(T, s));
    if (boot) I2 = ld_info8(row_info(T, s2));
    RowGather<LPR> rows;
    rows.issue(T, s2, boot);
    Scan sc;
    sc.dyn = 0u;
    sc.slow = 0u;

    // ---- own cell (branch-free over the first four entries)
    const uint32_t c1 = (I1.w[1] == epoch) ? I1.w[0] : 0u;
    {
        int pj = -1;
        uint32_t later = 0u;
        const uint32_t me = (uint32_t)a << 24;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t w = I1.w[4 + k];
            const bool same = ((uint32_t)k < c1) && ((w ^ me) < (1u << 24));
            const int j = (int)(w & kNone);
            pj = max(pj, (same && j < i) ? j : -1);
            later |= (same && j > i) ? 1u : 0u;
        }
        if (c1 > 4u) {  // crowded row: the rest of the inline entries, then the overflow list
            RowWriters w1;
            w1.iw = row_info(T, s);
            w1.e = make_uint4(0, 0, 0, 0);
            w1.count = c1;
            w1.ovf = (I1.w[3] == epoch) ? (I1.w[2] & kNone) : kNone;
            const uint32_t me2 = (uint32_t)a << 24;
            for_each_writer_from4(T, w1, [&](uint32_t w) {
                const bool same = (w ^ me2) < (1u << 24);
                const int j = (int)(w & kNone);
                pj = max(pj, (same && j < i) ? j : -1);
                later |= (same && j > i) ? 1u : 0u;
            });
        }
        sc.pj = pj;
        sc.later = later;
    }
    // ---- bootstrap row
    const uint32_t c2 = (I2.w[1] == epoch) ? I2.w[0] : 0u;  // 0 when !boot
    uint32_t contested = 0u;
    float m_fix = -INFINITY;
    sc.d0 = sc.d1 = sc.d2 = sc.d3 = -1;
    if (boot && mask2 == 0u) atomicOr(T.err, kErrEmpty);  // np.max of an empty selection (QLO:764)
#ifndef QE_FAST_WRITERS
#define QE_FAST_WRITERS 4u  // rows with at most this many writers take the branch-free path (0: always the general walk)
#endif
    if (c2 <= QE_FAST_WRITERS) {
        uint32_t ak[4], lk[4], beaten[4];
        int jk[4], key[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t w = I2.w[4 + k];
            ak[k] = (w >> 24) & 31u;
            jk[k] = (int)(w & kNone);
            lk[k] = ((uint32_t)k < c2) ? ((mask2 >> ak[k]) & 1u) : 0u;
            // priority among the writers of one cell: any earlier agent beats a later one, the latest earlier one
            // wins; cells written only by agents >= i keep one (arbitrary) representative
            key[k] = lk[k] ? ((jk[k] < i) ? (0x1000000 | jk[k]) : (3 - k)) : -1;
            contested |= lk[k] << ak[k];
            beaten[k] = 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int q = k + 1; q < 4; ++q) {
                const uint32_t same = lk[k] & lk[q] & (ak[k] == ak[q] ? 1u : 0u);
                const uint32_t kq = key[k] > key[q] ? 1u : 0u;
                beaten[q] |= same & kq;
                beaten[k] |= same & (kq ^ 1u);
            }
        int d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool rel = lk[k] && !beaten[k];
            const bool early = jk[k] < i;
            d[k] = (rel && early) ? jk[k] : -1;
            if (rel && !early) m_fix = fmax_plain(m_fix, __ldcg(T.tr_p + jk[k]));
        }
        sc.d0 = d[0]; sc.d1 = d[1]; sc.d2 = d[2]; sc.d3 = d[3];
    } else {  // crowded bootstrap row: one pass over all writers, latest earlier writer per action in best[]
        RowWriters w2;
        w2.iw = row_info(T, s2);
        w2.e = make_uint4(I2.w[4], I2.w[5], I2.w[6], I2.w[7]);
        w2.count = c2;
        w2.ovf = (I2.w[3] == epoch) ? (I2.w[2] & kNone) : kNone;
        // best[a] = latest earlier writer of (s', a) if bit a of `early` is set, else (bit a of `late`) some later writer
        uint32_t early = 0u, late = 0u;
        for_each_writer(T, w2, [&](uint32_t w) {
            const uint32_t aj = (w >> 24) & 31u, bit = 1u << aj;
            if (mask2 & bit) {
                const int j = (int)(w & kNone);
                int* b = best + aj * 256;
                if (j < i) {
                    *b = (early & bit) ? max(*b, j) : j;
                    early |= bit;
                } else if (!((early | late) & bit)) {
                    *b = j;
                    late |= bit;
                }
            }
        });
        contested = early | late;
        for (uint32_t b = late & ~early; b; b &= b - 1u)  // written only by agents >= i: Q0 as captured by one of them
            m_fix = fmax_plain(m_fix, __ldcg(T.tr_p + best[(__ffs(b) - 1) * 256]));
        if (SLOW_OK) {
            sc.dyn = early;  // resolved from best[] in the polling loop
        } else {             // up to four earlier writers fit the deferred record; more -> in-order pass
            int nd = 0;
            for (uint32_t b = early; b; b &= b - 1u) {
                const int jb = best[(__ffs(b) - 1) * 256];
                if (nd == 0) sc.d0 = jb; else if (nd == 1) sc.d1 = jb; else if (nd == 2) sc.d2 = jb; else if (nd == 3) sc.d3 = jb;
                else sc.slow = 1u;
                ++nd;
            }
        }
    }
    // cells of the bootstrap row nobody writes this step: straight from the table
    const float m_row = rows.row_max(boot ? (mask2 & ~contested) : 0u);
    sc.m = fmax_plain(m_row, m_fix);
    return sc;
}

__device__ __forceinline__ void publish_commit(const Table& T, int i, int s, int a, float v, uint32_t later, uint32_t epoch) {
    st_relaxed_u64(T.slot + i, ((uint64_t)epoch << 32) | (uint64_t)__float_as_uint(v));
    if (T.hold) T.later_buf[i] = (uint8_t)later;
    else if (!later) T.q[(size_t)s * T.ld + a] = v;  // last writer of the cell commits
}

// Deferred record (32 bytes per agent, written by the first pass for agents that have to wait for someone)
//   [0] m bits  [1] pj  [2..5] d0..d3  [6] a | term << 7 | later << 8 | slow << 9  [7] reward bits
__device__ __forceinline__ void save_deferred(const Table& T, int i, const Scan& sc, int a, bool term, float r) {
    U8 rec;
    rec.w[0] = __float_as_uint(sc.m);
    rec.w[1] = (uint32_t)sc.pj; rec.w[2] = (uint32_t)sc.d0; rec.w[3] = (uint32_t)sc.d1; rec.w[4] = (uint32_t)sc.d2; rec.w[5] = (uint32_t)sc.d3;
    rec.w[6] = (uint32_t)a | (term ? 0x80u : 0u) | (sc.later << 8) | (sc.slow << 9);
    rec.w[7] = __float_as_uint(r);
    st_u8(T.rec + (size_t)i * 8, rec);
}

// First pass over a tile of 32 agents: scan; agents that depend on nobody finish at once, the others are deferred.
// Returns 0 (finished), 1 (deferred) or 2 (deferred, crowded).  No waiting, no ordering requirement between tiles.
template <int LPR>
__device__ __forceinline__ int learn_first_pass(const Table& T, int* best, bool active, int i, int s, int a, float r, float p,
                                                int s2, bool term, uint32_t mask2, float lr, float gamma, uint32_t epoch) {
    const Scan sc = scan_writers<LPR, false>(T, best, active, i, s, a, s2, term, mask2, epoch);
    const bool free_now = (sc.pj & sc.d0 & sc.d1 & sc.d2 & sc.d3) < 0 && !sc.slow;  // all five are -1 <=> nothing to wait for
    if (active) {
        if (free_now) publish_commit(T, i, s, a, td_value(p, r, term ? 0.0f : sc.m, lr, gamma), sc.later, epoch);
        else save_deferred(T, i, sc, a, term, r);
    }
    return (active && !free_now) ? (sc.slow ? 2 : 1) : 0;
}
// per tile: which agents were deferred (dmask) and which of them need the in-order pass (smask)
__device__ __forceinline__ void store_tile_masks(const Table& T, int tile, int kind, int* nslow) {
    const uint32_t bd = __ballot_sync(kFull, kind == 1), bs = __ballot_sync(kFull, kind == 2);
    if ((threadIdx.x & 31) == 0) {
        T.dmask[tile] = bd;
        T.smask[tile] = bs;
        if (bs) atomicAdd(nslow, __popc(bs));
    }
}

// Enumerate the deferred agents of a group of tiles in increasing order.
// `mask` = this lane's tile mask (lane L <-> tile L of the group; 0 for lanes beyond the group).  Returns the number of deferred agents in the group;
// agent_of(rank) gives the agent index for rank < total.
struct DeferredGroup {
    uint32_t mask, incl;  // this lane's tile mask and the inclusive prefix sum of popc over lanes
    int total;
    __device__ __forceinline__ void init(uint32_t m) {
        mask = m;
        const int lane = threadIdx.x & 31;
        uint32_t x = (uint32_t)__popc(m);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, x, d);
            if (lane >= d) x += y;
        }
        incl = x;
        total = (int)__shfl_sync(kFull, x, 31);
    }
    // agent index (relative to the group's first agent) of the deferred agent with the given rank; all lanes call
    __device__ __forceinline__ int agent_of(int rank) const {
        int lo = 0;  // smallest lane t with incl[t] > rank
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const uint32_t v = __shfl_sync(kFull, incl, lo + step - 1);
            if ((int)v <= rank) lo += step;
        }
        lo = min(lo, 31);
        const uint32_t tm = __shfl_sync(kFull, mask, lo);
        const uint32_t ti = __shfl_sync(kFull, incl, lo);
        const int k = rank - (int)(ti - (uint32_t)__popc(tm));
        return lo * 32 + kth_set32(tm, max(k, 0));
    }
};

// ------------------------------------------------------------------ unfused select
template <int LPA>
__global__ void __launch_bounds__(256) select_kernel(Table T, const int32_t* __restrict__ states,
                                                     const uint32_t* __restrict__ mask_bits, Uniforms U, uint64_t thresh,
                                                     int deterministic, int empty_all, int32_t* __restrict__ actions, int n) {
    const uint32_t gm = group_mask<LPA>();
    const int l = threadIdx.x & (LPA - 1);
    const int groups = gridDim.x * (blockDim.x / LPA);
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    for (int i = blockIdx.x * (blockDim.x / LPA) + threadIdx.x / LPA; i < n; i += groups) {
        const int s = states[i];
        const uint32_t valid = mask_bits ? (mask_bits[i] & full) : full;
        const float4 v = ld_row4(T.q + (size_t)s * T.ld + 4 * l);
        const bool explore = !deterministic && ((uint64_t)U.draw(i, 0) < thresh);
        float qsa;
        const int a = select_group<LPA>(v, valid, T.A, explore, empty_all != 0, U.draw(i, 1), gm, &qsa);
        if (l == 0) actions[i] = a;
    }
}

// general action counts (A > 32 or byte masks): one thread per agent, three passes over the row
static __global__ void select_generic_kernel(Table T, const int32_t* __restrict__ states, const uint8_t* __restrict__ mask_bytes,
                                      Uniforms U, uint64_t thresh, int deterministic, int empty_all,
                                      int32_t* __restrict__ actions, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* row = T.q + (size_t)states[i] * T.ld;
    const uint8_t* mk = mask_bytes ? mask_bytes + (size_t)i * T.A : nullptr;
    const bool explore = !deterministic && ((uint64_t)U.draw(i, 0) < thresh);
    float best = -INFINITY;
    int nvalid = 0;
    for (int a = 0; a < T.A; ++a)
        if (!mk || mk[a]) { ++nvalid; best = fmax_plain(best, row[a]); }
    int cnt = 0;
    const bool all = (!explore && nvalid == 0 && empty_all);
    if (explore) cnt = nvalid;
    else if (all) cnt = T.A;
    else
        for (int a = 0; a < T.A; ++a) cnt += ((!mk || mk[a]) && row[a] == best);
    int res = -1;
    if (cnt > 0) {
        int idx = (int)pick(U.draw(i, 1), (uint32_t)cnt);
        for (int a = 0; a < T.A; ++a) {
            const bool c = all || ((!mk || mk[a]) && (explore || row[a] == best));
            if (c) { if (idx == 0) { res = a; break; } --idx; }
        }
    }
    actions[i] = res;
}

// ------------------------------------------------------------------ second pass (shared by both kernels)
#ifndef QE_FUSED_MIN_BLOCKS
#define QE_FUSED_MIN_BLOCKS 4
#endif
// Tiles per statically owned group of the sweep over deferred records: 8, or as many as it takes to give every warp
// of the grid ONE group when that is more (config 4's 4M agents on 4736 warps: 28 -- several groups per warp re-walk
// the tile masks and leave some warps with one group more than others: 4.0 -> 4.5 G agent-steps/s); at most 32 (one
// tile mask per lane).  Not fewer than 8: on config 3 (7 would do) the warps left without a group carry the in-order
// pass for the crowded agents, which is on the critical path -- spreading the sweeps over all warps cost 30 us a step.
__device__ __forceinline__ int fast_tiles(int n) {
    const int ntiles = (n + 31) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    return min(32, max(8, (ntiles + nwarps - 1) / nwarps));
}
constexpr int kSlowTiles = 16;  // tiles per dynamically claimed group of the in-order pass

// (a) Sweeps.  Every warp owns a fixed set of tile groups.  One sweep visits the still-deferred agents of those groups,
// 32 per batch: load the record, poll every predecessor once, finish the agent if all of them have published, clear
// its bit.  Nobody ever waits, so there is no ordering requirement; the DAG drains level by level.
template <int FT>  // FT > 0: group size known at compile time (the common 8); 0: ft_rt
__device__ __forceinline__ int sweep_deferred_impl(const Table& T, int n, const int32_t* cur, float lr, float gamma, uint32_t epoch, int ft_rt) {
    const int lane = threadIdx.x & 31;
    const int kFastTiles = FT > 0 ? FT : ft_rt;
    const int ntiles = (n + 31) >> 5, ngroups = (ntiles + kFastTiles - 1) / kFastTiles;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    int left = 0;  // agents of this warp still deferred after the sweep
    for (int g = warp; g < ngroups; g += nwarps) {
        const int tile = g * kFastTiles + lane;
        const bool mine = lane < kFastTiles && tile < ntiles;
        const uint32_t mask = mine ? __ldcg(T.dmask + tile) : 0u;
        DeferredGroup grp;
        grp.init(mask);
        if (grp.total == 0) continue;
        uint32_t keep = mask;
        for (int b = 0; b < grp.total; b += 32) {
            const int rank = b + lane;
            const bool act = rank < grp.total;
            const int rel = grp.agent_of(act ? rank : 0);
            bool fin = false;
            if (act) {
                const int i = g * (kFastTiles * 32) + rel;
                const U8 rec = ld_info8(T.rec + (size_t)i * 8);
                const int pj = (int)rec.w[1], d0 = (int)rec.w[2], d1 = (int)rec.w[3], d2 = (int)rec.w[4], d3 = (int)rec.w[5];
                const uint64_t wp = pj >= 0 ? ld_relaxed_u64(T.slot + pj) : 0ull;
                const uint64_t w0 = d0 >= 0 ? ld_relaxed_u64(T.slot + d0) : 0ull;
                const uint64_t w1 = d1 >= 0 ? ld_relaxed_u64(T.slot + d1) : 0ull;
                const uint64_t w2 = d2 >= 0 ? ld_relaxed_u64(T.slot + d2) : 0ull;
                const uint64_t w3 = d3 >= 0 ? ld_relaxed_u64(T.slot + d3) : 0ull;
                const bool ok = (pj < 0 || (uint32_t)(wp >> 32) == epoch) && (d0 < 0 || (uint32_t)(w0 >> 32) == epoch) &&
                                (d1 < 0 || (uint32_t)(w1 >> 32) == epoch) && (d2 < 0 || (uint32_t)(w2 >> 32) == epoch) &&
                                (d3 < 0 || (uint32_t)(w3 >> 32) == epoch);
                if (ok) {
                    float m = __uint_as_float(rec.w[0]);
                    if (d0 >= 0) m = fmax_plain(m, __uint_as_float((uint32_t)w0));
                    if (d1 >= 0) m = fmax_plain(m, __uint_as_float((uint32_t)w1));
                    if (d2 >= 0) m = fmax_plain(m, __uint_as_float((uint32_t)w2));
                    if (d3 >= 0) m = fmax_plain(m, __uint_as_float((uint32_t)w3));
                    const float pe = pj >= 0 ? __uint_as_float((uint32_t)wp) : __ldcg(T.tr_p + i);
                    const bool term = (rec.w[6] & 0x80u) != 0u;
                    const float v = td_value(pe, __uint_as_float(rec.w[7]), term ? 0.0f : m, lr, gamma);
                    publish_commit(T, i, __ldcg(cur + i), (int)(rec.w[6] & 0x7Fu), v, (rec.w[6] >> 8) & 1u, epoch);
                    fin = true;
                }
            }
#pragma unroll
            for (int t = 0; t < kFastTiles; ++t) {
                const uint32_t bits = __reduce_or_sync(kFull, (fin && (rel >> 5) == t) ? (1u << (rel & 31)) : 0u);
                if (lane == t) keep &= ~bits;
            }
        }
        if (mine && keep != mask) __stcg(T.dmask + tile, keep);
        left += (int)__reduce_add_sync(kFull, (uint32_t)__popc(keep));
    }
    return left;
}

template <typename Dummy = void>
__device__ __forceinline__ int sweep_deferred(const Table& T, int n, const int32_t* cur, float lr, float gamma, uint32_t epoch) {
    const int ft = fast_tiles(n);
    return ft == 8 ? sweep_deferred_impl<8>(T, n, cur, lr, gamma, epoch, 8) : sweep_deferred_impl<0>(T, n, cur, lr, gamma, epoch, ft);
}

// Once at most 32 deferred agents are left to a warp they move into its lanes and are polled back to back (the hop
// from a predecessor's publication to the dependant's own costs one L2 round trip instead of one sweep).
struct ResidentLanes {
    int i, pj, d0, d1, d2, d3;
    float m, r;
    uint32_t flags;
    bool busy;
    __device__ __forceinline__ void load(const Table& T, int n) {
        const int lane = threadIdx.x & 31;
        const int kFastTiles = fast_tiles(n);
        const int ntiles = (n + 31) >> 5, ngroups = (ntiles + kFastTiles - 1) / kFastTiles;
        const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
        busy = false;
        i = 0;
        int have = 0;
        for (int g = warp; g < ngroups; g += nwarps) {
            const int tile = g * kFastTiles + lane;
            DeferredGroup grp;
            grp.init((lane < kFastTiles && tile < ntiles) ? __ldcg(T.dmask + tile) : 0u);
            if (grp.total == 0) continue;
            const int rank = lane - have;
            const bool act = rank >= 0 && rank < grp.total;
            const int rel = grp.agent_of(act ? rank : 0);
            if (act) { i = g * (kFastTiles * 32) + rel; busy = true; }
            have += grp.total;
        }
        pj = d0 = d1 = d2 = d3 = -1;
        m = r = 0.0f;
        flags = 0u;
        if (busy) {
            const U8 rec = ld_info8(T.rec + (size_t)i * 8);
            m = __uint_as_float(rec.w[0]);
            pj = (int)rec.w[1]; d0 = (int)rec.w[2]; d1 = (int)rec.w[3]; d2 = (int)rec.w[4]; d3 = (int)rec.w[5];
            flags = rec.w[6];
            r = __uint_as_float(rec.w[7]);
        }
    }
    // one poll of every lane's outstanding predecessors; returns true while some lane is still waiting
    __device__ __forceinline__ bool step(const Table& T, const int32_t* cur, float lr, float gamma, uint32_t epoch) {
        if (busy) {
            const uint64_t wp = pj >= 0 ? ld_relaxed_u64(T.slot + pj) : 0ull;
            const uint64_t w0 = d0 >= 0 ? ld_relaxed_u64(T.slot + d0) : 0ull;
            const uint64_t w1 = d1 >= 0 ? ld_relaxed_u64(T.slot + d1) : 0ull;
            const uint64_t w2 = d2 >= 0 ? ld_relaxed_u64(T.slot + d2) : 0ull;
            const uint64_t w3 = d3 >= 0 ? ld_relaxed_u64(T.slot + d3) : 0ull;
            const bool ok = (pj < 0 || (uint32_t)(wp >> 32) == epoch) && (d0 < 0 || (uint32_t)(w0 >> 32) == epoch) &&
                            (d1 < 0 || (uint32_t)(w1 >> 32) == epoch) && (d2 < 0 || (uint32_t)(w2 >> 32) == epoch) &&
                            (d3 < 0 || (uint32_t)(w3 >> 32) == epoch);
            if (ok) {
                float mm = m;
                if (d0 >= 0) mm = fmax_plain(mm, __uint_as_float((uint32_t)w0));
                if (d1 >= 0) mm = fmax_plain(mm, __uint_as_float((uint32_t)w1));
                if (d2 >= 0) mm = fmax_plain(mm, __uint_as_float((uint32_t)w2));
                if (d3 >= 0) mm = fmax_plain(mm, __uint_as_float((uint32_t)w3));
                const float pe = pj >= 0 ? __uint_as_float((uint32_t)wp) : __ldcg(T.tr_p + i);
                const bool term = (flags & 0x80u) != 0u;
                const float v = td_value(pe, r, term ? 0.0f : mm, lr, gamma);
                publish_commit(T, i, __ldcg(cur + i), (int)(flags & 0x7Fu), v, (flags >> 8) & 1u, epoch);
                busy = false;
            }
        }
        return __any_sync(kFull, busy);
    }
};

// (b) In-order pass for crowded agents (state in registers + best[] in shared memory).  They are taken in increasing
// agent order: groups of tiles are claimed from a global cursor and inside a group the agents are handed to free lanes
// by rank.  A lane keeps its agent until every predecessor has published (one poll per step()), then takes the next
// one.  Progress: the smallest unfinished agent has no unfinished predecessor; if it is crowded it is in some lane
// already or next in line for a warp whose lanes all hold smaller (hence finished) agents; otherwise the next sweep
// of its owner finishes it.  All warps are co-resident (cooperative launch) and no warp ever blocks.
template <int LPR>
struct InOrderLanes {
    DeferredGroup grp;
    int g, next_rank;
    bool more, busy, term;
    int i, s, a, pj, d0, d1, d2, d3;
    float r, pe, m;
    uint32_t later, dyn;
    __device__ __forceinline__ void init(bool any_slow) {
        grp.total = 0; grp.mask = 0u; grp.incl = 0u;
        g = -1; next_rank = 0;
        more = any_slow; busy = false; term = false;
        i = s = a = 0; pj = d0 = d1 = d2 = d3 = -1;
        r = pe = m = 0.0f; later = dyn = 0u;
    }
    // one refill + one poll; returns false once this warp has nothing in flight and no group is left
    template <typename MaskFn>
    __device__ __forceinline__ bool step(const Table& T, int* best, int* cursor, int n, const int32_t* cur, const int32_t* nxt,
                                         MaskFn mask_of, float lr, float gamma, uint32_t epoch) {
        const int lane = threadIdx.x & 31;
        const int ntiles = (n + 31) >> 5, ngroups = (ntiles + kSlowTiles - 1) / kSlowTiles;
        const uint32_t freeb = __ballot_sync(kFull, !busy);
        if (freeb) {
            while (more && next_rank >= grp.total) {  // group exhausted: claim the next one
                if (lane == 0) g = atomicAdd(cursor, 1);
                g = __shfl_sync(kFull, g, 0);
                more = g < ngroups;
                next_rank = 0;
                grp.total = 0;
                if (more) {
                    const int tile = g * kSlowTiles + lane;
                    grp.init((lane < kSlowTiles && tile < ntiles) ? __ldcg(T.smask + tile) : 0u);
                }
            }
            if (next_rank < grp.total) {
                const int rank = next_rank + __popc(freeb & ((1u << lane) - 1u));
                const bool take = !busy && rank < grp.total;
                const int idx = g * (kSlowTiles * 32) + grp.agent_of(take ? rank : 0);
                next_rank = min(next_rank + __popc(freeb), grp.total);
                int s2 = 0;
                uint32_t m2 = 0u;
                if (take) {
                    i = idx;
                    const U8 rec = ld_info8(T.rec + (size_t)i * 8);
                    s = __ldcg(cur + i);
                    pe = __ldcg(T.tr_p + i);
                    a = (int)(rec.w[6] & 0x7Fu);
                    term = (rec.w[6] & 0x80u) != 0u;
                    r = __uint_as_float(rec.w[7]);
                    s2 = __ldcg(nxt + i);
                    m2 = mask_of(i, s2);
                    busy = true;
                }
                // full scan (all lanes call: transposed gathers inside); best[] stays valid while the lane is busy
                const Scan sc = scan_writers<LPR, true>(T, best, take, i, s, a, s2, term, m2, epoch);
                if (take) { m = sc.m; pj = sc.pj; d0 = sc.d0; d1 = sc.d1; d2 = sc.d2; d3 = sc.d3; later = sc.later; dyn = sc.dyn; }
            } else if (freeb == kFull && !more) {
                return false;  // every lane is free and nothing is left
            }
        }
        if (busy) {  // every outstanding poll is issued before any is examined
            const uint64_t wp = pj >= 0 ? ld_relaxed_u64(T.slot + pj) : 0ull;
            const uint64_t w0 = d0 >= 0 ? ld_relaxed_u64(T.slot + d0) : 0ull;
            const uint64_t w1 = d1 >= 0 ? ld_relaxed_u64(T.slot + d1) : 0ull;
            const uint64_t w2 = d2 >= 0 ? ld_relaxed_u64(T.slot + d2) : 0ull;
            const uint64_t w3 = d3 >= 0 ? ld_relaxed_u64(T.slot + d3) : 0ull;
            if (pj >= 0 && (uint32_t)(wp >> 32) == epoch) { pe = __uint_as_float((uint32_t)wp); pj = -1; }
            if (d0 >= 0 && (uint32_t)(w0 >> 32) == epoch) { m = fmax_plain(m, __uint_as_float((uint32_t)w0)); d0 = -1; }
            if (d1 >= 0 && (uint32_t)(w1 >> 32) == epoch) { m = fmax_plain(m, __uint_as_float((uint32_t)w1)); d1 = -1; }
            if (d2 >= 0 && (uint32_t)(w2 >> 32) == epoch) { m = fmax_plain(m, __uint_as_float((uint32_t)w2)); d2 = -1; }
            if (d3 >= 0 && (uint32_t)(w3 >> 32) == epoch) { m = fmax_plain(m, __uint_as_float((uint32_t)w3)); d3 = -1; }
            if (dyn) {  // crowded bootstrap row
                uint32_t left = 0;
                for (uint32_t b = dyn; b; b &= b - 1u) {
                    const int a2 = __ffs(b) - 1;
                    const uint64_t w = ld_relaxed_u64(T.slot + best[a2 * 256]);
                    if ((uint32_t)(w >> 32) == epoch) m = fmax_plain(m, __uint_as_float((uint32_t)w));
                    else left |= 1u << a2;
                }
                dyn = left;
            }
            if ((pj & d0 & d1 & d2 & d3) < 0 && !dyn) {
                publish_commit(T, i, s, a, td_value(pe, r, term ? 0.0f : m, lr, gamma), later, epoch);
                busy = false;
            }
        }
        return true;
    }
};

template <int LPR, typename MaskFn>
__device__ __forceinline__ void second_pass_all(const Table& T, int* best, int* cursor, bool any_slow, int n, const int32_t* cur,
                                                const int32_t* nxt, MaskFn mask_of, float lr, float gamma, uint32_t epoch) {
    InOrderLanes<LPR> io;
    io.init(any_slow);
    ResidentLanes res;
    bool fast_left = true, slow_left = any_slow, resident = false;
    const uint64_t t0 = global_ns();
    for (uint32_t spins = 0; fast_left || slow_left; ++spins) {
        if (fast_left) {
            if (!resident) {
                const int pending = sweep_deferred(T, n, cur, lr, gamma, epoch);
                if (pending == 0) fast_left = false;
                else if (pending <= 32) { res.load(T, n); resident = true; }
            } else {
                fast_left = res.step(T, cur, lr, gamma, epoch);
            }
        }
        if (slow_left) slow_left = io.step(T, best, cursor, n, cur, nxt, mask_of, lr, gamma, epoch);
        if ((spins & 255u) == 255u && global_ns() - t0 > kTimeoutNs) { atomicOr(T.err, kErrTimeout); break; }
    }
}

// ------------------------------------------------------------------ unfused exact learn (cooperative)
template <int LPR>
__global__ void __launch_bounds__(256) learn_exact_kernel(Table T, const int32_t* __restrict__ states,
                                                          const int32_t* __restrict__ actions, const float* __restrict__ rewards,
                                                          const int32_t* __restrict__ next_states,
                                                          const uint8_t* __restrict__ terminated,
                                                          const uint32_t* __restrict__ next_mask_bits, float lr, float gamma,
                                                          uint32_t epoch, int* cursor, int n) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_best[8 * LPR * 256];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    for (int i = tid; i < n; i += nthreads) {
        const int s = states[i], a = actions[i];
        row_insert(T, i, s, a, __ldcg(T.q + (size_t)s * T.ld + a), epoch);
    }
    grid.sync();
    {
        int* claim = cursor + 4;
        int base = 0;
        if (lane == 0) base = atomicAdd(claim, 32);
        base = __shfl_sync(kFull, base, 0);
        while (base < n) {
            int next_base = 0;
            if (lane == 0) next_base = atomicAdd(claim, 32);
            const int i = base + lane;
            const bool active = i < n;
            const int ii = active ? i : 0;
            const uint32_t m2 = next_mask_bits ? (next_mask_bits[ii] & full) : full;
            const int kind = learn_first_pass<LPR>(T, s_best + threadIdx.x, active, i, states[ii], actions[ii], rewards[ii],
                                                   __ldcg(T.tr_p + ii), next_states[ii], terminated[ii] != 0, m2, lr, gamma, epoch);
            store_tile_masks(T, base >> 5, kind, cursor + 2);
            base = __shfl_sync(kFull, next_base, 0);
        }
    }
    grid.sync();
    second_pass_all<LPR>(T, s_best + threadIdx.x, cursor, __ldcg(cursor + 2) != 0, n, states, next_states,
                         [&](int i, int) { return next_mask_bits ? (__ldg(next_mask_bits + i) & full) : full; }, lr, gamma, epoch);
}

// general fallback (A > 32 / byte masks): the reference loop itself, one warp walks the agents in order
static __global__ void learn_sequential_kernel(Table T, const int32_t* __restrict__ states, const int32_t* __restrict__ actions,
                                        const float* __restrict__ rewards, const int32_t* __restrict__ next_states,
                                        const uint8_t* __restrict__ terminated, const uint8_t* __restrict__ next_mask_bytes,
                                        const uint32_t* __restrict__ next_mask_bits, float lr, float gamma, int n) {
    const int lane = threadIdx.x;
    for (int i = 0; i < n; ++i) {
        float m = 0.0f;
        if (!terminated[i]) {
            const volatile float* row = T.q + (size_t)next_states[i] * T.ld;
            m = -INFINITY;
            for (int a = lane; a < T.A; a += 32) {
                const bool ok = next_mask_bytes ? next_mask_bytes[(size_t)i * T.A + a] != 0
                                                : (next_mask_bits ? ((next_mask_bits[i] >> a) & 1u) != 0 : true);
                if (ok) m = fmax_plain(m, row[a]);
            }
            for (int d = 16; d > 0; d >>= 1) m = fmax_plain(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
            if (m == -INFINITY && lane == 0) atomicOr(T.err, kErrEmpty);
        }
        if (lane == 0) {
            volatile float* cell = T.q + (size_t)states[i] * T.ld + actions[i];
            *cell = td_value(*cell, rewards[i], m, lr, gamma);
        }
        __syncwarp();
    }
}

// accumulate mode (learn_vec, QLO:853-891): deltas from the snapshot, then atomic scatter-add
static __global__ void learn_delta_kernel(Table T, const int32_t* __restrict__ states, const int32_t* __restrict__ actions,
                                   const float* __restrict__ rewards, const int32_t* __restrict__ next_states,
                                   const uint8_t* __restrict__ terminated, const uint8_t* __restrict__ next_mask_bytes,
                                   const uint32_t* __restrict__ next_mask_bits, float lr, float gamma, float* __restrict__ delta,
                                   int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* row = T.q + (size_t)next_states[i] * T.ld;
    float m = -INFINITY;
    for (int a = 0; a < T.A; ++a) {
        const bool ok = next_mask_bytes ? next_mask_bytes[(size_t)i * T.A + a] != 0
                                        : (next_mask_bits ? ((next_mask_bits[i] >> a) & 1u) != 0 : true);
        if (ok) m = fmax_plain(m, row[a]);
    }
    // targets = r + gamma * max * (1 - terminated)
    const float target = rewards[i] + gamma * m * (terminated[i] ? 0.0f : 1.0f);
    delta[i] = lr * (target - T.q[(size_t)states[i] * T.ld + actions[i]]);
}
static __global__ void learn_scatter_kernel(Table T, const int32_t* __restrict__ states, const int32_t* __restrict__ actions,
                                     const float* __restrict__ delta, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(T.q + (size_t)states[i] * T.ld + actions[i], delta[i]);
}

static __global__ void gather_kernel(Table T, const int32_t* __restrict__ states, const int32_t* __restrict__ actions,
                              float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = T.q[(size_t)states[i] * T.ld + actions[i]];
}

static __global__ void gather_rows_kernel(Table T, const int32_t* __restrict__ states, float* __restrict__ out, int n) {
    const size_t total = (size_t)n * T.A;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        const size_t i = x / T.A;
        out[x] = T.q[(size_t)states[i] * T.ld + (x - i * T.A)];
    }
}

static __global__ void table_fill_kernel(Table T, int64_t S, float value, uint32_t seed, int random, uint64_t state_base) {
    const size_t total = (size_t)S * T.ld;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        const int a = (int)(x % T.ld);
        const size_t s = x / T.ld + state_base;  // global state id: a shard fills exactly its slice of the whole table
        float v = 0.0f;
        if (a < T.A) v = random ? (float)(fmix32((uint32_t)(s * (size_t)T.A + a) ^ (seed * kGold)) >> 8) * 5.9604644775390625e-08f : value;
        T.q[x] = v;
    }
}

// ------------------------------------------------------------------ sharded / replicated table helpers
// commit of a held TD update: the last writer of every cell stores its published value
static __global__ void learn_commit_kernel(Table T, const int32_t* __restrict__ states, const int32_t* __restrict__ actions, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !T.later_buf[i]) T.q[(size_t)states[i] * T.ld + actions[i]] = __uint_as_float((uint32_t)ld_relaxed_u64(T.slot + i));
}
// Bootstrap requests of agents that live on another shard: m = max over the legal actions of the value of
// (row, a') just before an agent that sorts before local agent `pos` (all local agents < pos are earlier, all others
// later).  Values of earlier writers come from their published slots (the held update of this epoch), everything
// else from the untouched table.  use_versions = 0: plain snapshot max.
static __global__ void serve_bootstrap_kernel(Table T, const int32_t* __restrict__ rows, const int32_t* __restrict__ pos,
                                       const uint32_t* __restrict__ masks, float* __restrict__ out, int n, uint32_t epoch,
                                       int use_versions) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int row = rows[r], before = pos[r];
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    const uint32_t legal = masks ? (masks[r] & full) : full;
    RowWriters w = load_writers(T, row, epoch);
    if (!use_versions) w.count = 0;
    // the Q row, one 32-byte sector per load (at most four: A <= 32)
    const int sectors = (T.A + 7) >> 3;
    float m = -INFINITY;
    uint32_t versioned = 0u;  // legal actions that some earlier local agent writes
    if (w.count) {
        for (uint32_t b = legal; b; b &= b - 1u) {
            const int a2 = __ffs(b) - 1;
            int best = -1;
            for_each_writer(T, w, [&](uint32_t e) {
                const int j = (int)(e & kNone);
                if ((int)(e >> 24) == a2 && j < before) best = max(best, j);
            });
            if (best >= 0) {
                versioned |= 1u << a2;
                m = fmax_plain(m, __uint_as_float((uint32_t)ld_relaxed_u64(T.slot + best)));
            }
        }
    }
    for (int q = 0; q < sectors; ++q) {
        const uint32_t my = ((legal & ~versioned) >> (8 * q)) & 0xFFu;
        if (my) m = fmax_plain(m, max8(ld_row8(T.q + (size_t)row * T.ld + 8 * q), my));
    }
    out[r] = m;
}
// replicated table: delta[s][a] = Q[s][a] - base[s][a]   (dense [S][A] buffers)
static __global__ void table_delta_kernel(Table T, int64_t S, const float* __restrict__ base, float* __restrict__ delta) {
    const size_t total = (size_t)S * T.A, stride = (size_t)gridDim.x * blockDim.x;
    if (T.ld == T.A) {  // rows without padding (A = 8, 16, 32): the table IS the dense array, 16 bytes per lane and access
        const float4* q4 = reinterpret_cast<const float4*>(T.q);
        const float4* b4 = reinterpret_cast<const float4*>(base);
        float4* d4 = reinterpret_cast<float4*>(delta);
        for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total / 4; x += stride) {
            const float4 a = q4[x], b = b4[x];
            d4[x] = make_float4(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z), __fsub_rn(a.w, b.w));
        }
        return;
    }
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        const size_t s = x / T.A;
        delta[x] = __fsub_rn(T.q[s * T.ld + (x - s * T.A)], base[x]);
    }
}
// ... and Q = base = base + sum of the ranks' deltas
static __global__ void table_merge_kernel(Table T, int64_t S, float* __restrict__ base, const float* __restrict__ delta_sum) {
    const size_t total = (size_t)S * T.A, stride = (size_t)gridDim.x * blockDim.x;
    if (T.ld == T.A) {
        float4* q4 = reinterpret_cast<float4*>(T.q);
        float4* b4 = reinterpret_cast<float4*>(base);
        const float4* d4 = reinterpret_cast<const float4*>(delta_sum);
        for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total / 4; x += stride) {
            const float4 b = b4[x], d = d4[x];
            const float4 v = make_float4(__fadd_rn(b.x, d.x), __fadd_rn(b.y, d.y), __fadd_rn(b.z, d.z), __fadd_rn(b.w, d.w));
            b4[x] = v;
            q4[x] = v;
        }
        return;
    }
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        const size_t s = x / T.A;
        const float v = __fadd_rn(base[x], delta_sum[x]);
        base[x] = v;
        T.q[s * T.ld + (x - s * T.A)] = v;
    }
}
static __global__ void table_export_kernel(Table T, int64_t S, float* __restrict__ dense) {
    const size_t total = (size_t)S * T.A, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        const size_t s = x / T.A;
        dense[x] = T.q[s * T.ld + (x - s * T.A)];
    }
}
static __global__ void table_import_kernel(Table T, int64_t S, const float* __restrict__ dense) {
    const size_t total = (size_t)S * T.A, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        const size_t s = x / T.A;
        T.q[s * T.ld + (x - s * T.A)] = dense[x];
    }
}

// ------------------------------------------------------------------ unfused environments (one thread per agent)
static __global__ void ttt_reset_kernel(uint32_t* __restrict__ boards, int32_t* __restrict__ states, uint32_t* __restrict__ masks,
                                 Uniforms U, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = ttt_reset(U.draw(i, 3), U.draw(i, 4));
    boards[i] = b;
    states[i] = ttt_state(b & 0x3FFFFu);
    masks[i] = ttt_empties(b);
}
static __global__ void ttt_step_kernel(uint32_t* __restrict__ boards, const int32_t* __restrict__ actions, Uniforms U,
                                int32_t* __restrict__ next_states, uint32_t* __restrict__ next_masks, float* __restrict__ rewards,
                                uint8_t* __restrict__ terminated, int* err, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t b = boards[i];
    float r = 0.0f;
    bool term = false;
    if (!ttt_step(b, actions[i], U.draw(i, 2), U.draw(i, 3), U.draw(i, 4), r, term)) { atomicOr(err, kErrInvalidMove); return; }
    boards[i] = b;
    next_states[i] = ttt_state(b & 0x3FFFFu);
    next_masks[i] = ttt_empties(b);
    rewards[i] = r;
    terminated[i] = term;
}
static __global__ void mdp_reset_kernel(int32_t* __restrict__ states, uint32_t* __restrict__ masks, uint32_t S, int A, uint32_t env_seed,
                                 Uniforms U, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t s = (int32_t)pick(U.draw(i, 3), S);
    states[i] = s;
    if (masks) masks[i] = mdp_mask((uint32_t)s, A, env_seed);
}
static __global__ void mdp_step_kernel(int32_t* __restrict__ states, const int32_t* __restrict__ actions, uint32_t S, int A,
                                uint32_t env_seed, uint64_t term_thresh, Uniforms U, uint32_t* __restrict__ next_masks,
                                float* __restrict__ rewards, uint8_t* __restrict__ terminated, int* err, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = actions[i];
    if (a < 0 || a >= A) { atomicOr(err, kErrInvalidMove); return; }
    int32_t s = states[i];
    float r;
    bool term;
    mdp_step(s, a, S, A, env_seed, term_thresh, U.draw(i, 2), U.draw(i, 3), r, term);
    states[i] = s;
    if (next_masks) next_masks[i] = mdp_mask((uint32_t)s, A, env_seed);
    rewards[i] = r;
    terminated[i] = term;
}

static __global__ void mdp_masks_kernel(const int32_t* __restrict__ states, uint32_t* __restrict__ masks, int A, uint32_t env_seed, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) masks[i] = mdp_mask((uint32_t)states[i], A, env_seed);
}

// ------------------------------------------------------------------ fused persistent loop
struct FusedArgs {
    int env_kind, n, steps;
    int32_t* st_a;   // states (current at even local steps)
    int32_t* st_b;   // scratch
    uint32_t* envw;
    float* ep_ret;
    uint8_t* tr_a;   // [cap] action | term << 7
    float* tr_r;     // [cap]
    uint32_t env_seed, episode_len;
    uint64_t term_thresh;
    int64_t S;
    const uint64_t* eps_thresh;  // device [steps]
    const float* lr;             // device [steps]
    const uint32_t* uniforms;
    int slots;
    uint32_t stream_seed, t0, agent0, env_stream_seed, env_t0;
    int empty_all, use_masks;
    float gamma;
    uint32_t step0;              // engine-global step counter at launch (epoch source)
    int* tile_counter;           // [6] in-order cursors [0,1], crowded-agent counts [2,3], first-pass tile claims [4,5] (double-buffered across steps), 0 at launch
    int32_t* trace_actions;
    float* trace_rewards;
    uint8_t* trace_term;
    int32_t* trace_next;
    float* trace_epret;
    double* ep_sum;
    unsigned long long* ep_count;
    int evaluate;                // 1: no TD update (evaluation loops, BRT:293-384): select + env step only
    int accumulate;              // 1: learn_vec semantics (QLO:819-891): bootstrap from the table as it was before the step, atomicAdd scatter
    uint64_t* phase_ns;          // optional [31]: %globaltimer at launch and after each of the 3 phases of the first 10 steps
};

template <int ENV>
__device__ __forceinline__ uint32_t env_mask(int s, uint32_t envw, int A, uint32_t env_seed) {
    if (ENV == 0) return mdp_mask((uint32_t)s, A, env_seed);
    if (ENV == 1) return ttt_empties(envw);
    return A >= 32 ? 0xFFFFFFFFu : ((1u << A) - 1u);
}

template <int ENV, int LPR, bool ACC = false>  // ACC: the plain-atomics (learn_vec) update instead of the exact sequential one
__global__ void __launch_bounds__(256, QE_FUSED_MIN_BLOCKS) fused_kernel(Table T, FusedArgs F) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_sum[8];
    __shared__ unsigned int s_cnt[8];
    __shared__ int s_best[8 * LPR * 256];
    const int lane = threadIdx.x & 31;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    const int n = F.n;
    const bool clk = F.phase_ns != nullptr && tid == 0;
    if (clk) F.phase_ns[0] = global_ns();

    for (int k = 0; k < F.steps; ++k) {
        const uint32_t epoch = F.step0 + (uint32_t)k + 1u;
        int32_t* cur = (k & 1) ? F.st_b : F.st_a;
        int32_t* nxt = (k & 1) ? F.st_a : F.st_b;
        Uniforms U{F.uniforms ? F.uniforms + (size_t)k * n * F.slots : nullptr, F.slots, F.stream_seed, F.t0 + (uint32_t)k, F.agent0,
                   F.env_stream_seed, F.env_t0 + (uint32_t)k};
        const uint64_t thresh = F.eps_thresh[k];
        const float lr = F.lr[k];
        double loc_sum = 0.0;
        unsigned int loc_cnt = 0;

        // ---------------- phase A: select + env step + writer registration (one lane per agent)
        for (int base = (tid & ~31); base < n; base += nthreads) {
            const int i = base + lane;
            const bool active = i < n;
            int s = 0;
            uint32_t ew = 0u, valid = 0u, bits1 = 0u, ticket = 0u;
            bool explore = false;
            unsigned long long* info = nullptr;
            if (active) {
                s = cur[i];
                // register as a writer of row s right away: the ticket travels while the row is fetched
                info = reinterpret_cast<unsigned long long*>(row_info(T, s));
                if (!F.evaluate && !ACC) {
                    atomicMax(info, (unsigned long long)epoch << 32);  // a stale (older-epoch) counter restarts at {epoch, 0}
                    ticket = (uint32_t)atomicAdd(info, 1ull);
                }
                if (ENV != 0) ew = F.envw[i];
                valid = F.use_masks ? env_mask<ENV>(s, ew, T.A, F.env_seed) : full;
                explore = (uint64_t)U.draw(i, 0) < thresh;
                bits1 = U.draw(i, 1);
            }
            RowGather<LPR> rows;
            rows.issue(T, s, active);
            float mx;
            uint32_t tie;
            rows.row_max_tie(valid, mx, tie);
            int a = pick_action(T.A, valid, tie, explore, F.empty_all != 0, bits1);
            if (active && a < 0) { atomicOr(T.err, kErrEmpty); a = 0; }  // reported by qe_sync; the run's results are void
            a = max(a, 0);
            const float p = rows.value_of(a);
            if (active) {
                int32_t s2 = s;
                float r = 0.0f;
                bool term = false;
                if (ENV == 0) mdp_step(s2, a, (uint32_t)F.S, T.A, F.env_seed, F.term_thresh, U.draw(i, 2), U.draw(i, 3), r, term);
                else if (ENV == 1) {
                    if (!ttt_step(ew, a, U.draw(i, 2), U.draw(i, 3), U.draw(i, 4), r, term)) atomicOr(T.err, kErrInvalidMove);
                    s2 = ttt_state(ew & 0x3FFFFu);
                } else {  // bandit: reward = action, terminate every episode_len steps (rigged_two_armed_bandit.py:71-80)
                    r = (float)a;
                    ew += 1u;
                    term = ew >= F.episode_len;
                    if (term) ew = 0u;
                    s2 = 0;
                }
                // writer entry {agent, action}: inline in the row block, then the row's spill array, then the overflow list
                if (!F.evaluate && !ACC) store_entry(T, info, ticket, i, a, epoch);
                T.tr_p[i] = p;
                nxt[i] = s2;
                if (ENV != 0) F.envw[i] = ew;
                F.tr_a[i] = (uint8_t)(a | (term ? 0x80 : 0));
                F.tr_r[i] = r;
                float acc = F.ep_ret[i] + r;  // BRT:212
                float fin = __int_as_float(0x7FC00000);
                if (term) { fin = acc; loc_sum += (double)acc; ++loc_cnt; acc = 0.0f; }  // BRT:218-221
                F.ep_ret[i] = acc;
                const size_t o = (size_t)k * n + i;
                if (F.trace_actions) F.trace_actions[o] = a;
                if (F.trace_rewards) F.trace_rewards[o] = r;
                if (F.trace_term) F.trace_term[o] = term;
                if (F.trace_next) F.trace_next[o] = s2;
                if (F.trace_epret) F.trace_epret[o] = fin;
            }
            __syncwarp();
        }
        if (F.ep_count) {  // block-level reduction of the episode statistics, one atomic per block
            for (int d = 16; d > 0; d >>= 1) {
                loc_sum += __shfl_xor_sync(kFull, loc_sum, d);
                loc_cnt += __shfl_xor_sync(kFull, loc_cnt, d);
            }
            if (lane == 0) { s_sum[threadIdx.x >> 5] = loc_sum; s_cnt[threadIdx.x >> 5] = loc_cnt; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double bs = 0.0;
                unsigned int bc = 0;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { bs += s_sum[w]; bc += s_cnt[w]; }
                if (bc) { atomicAdd(F.ep_sum, bs); atomicAdd(F.ep_count, (unsigned long long)bc); }
            }
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[1 + 3 * k] = global_ns();
        if (F.evaluate) continue;  // the table is read-only: nothing to update, the barrier above orders the state buffers
        if (ACC) {
            // ---------------- learn_vec (QLO:819-891): every bootstrap and every prediction comes from the table as it
            // was before this step (the prediction was captured in phase A), the increments are scattered with atomicAdd
            // (np.add.at accumulates in agent order; atomics in any order -> equal up to fp32 rounding of the sum)
            for (int base = (tid & ~31); base < n; base += nthreads) {
                const int i = base + lane;
                const bool active = i < n;
                const int ii = active ? i : 0;
                const uint8_t at = F.tr_a[ii];
                const bool boot = active && !(at & 0x80);
                const int s2 = nxt[ii];
                const uint32_t ew = (ENV == 1) ? F.envw[ii] : 0u;
                const uint32_t m2 = F.use_masks ? env_mask<ENV>(s2, ew, T.A, F.env_seed) : full;
                RowGather<LPR> rows;
                rows.issue(T, s2, boot);
                const float m = rows.row_max(boot ? m2 : 0u);
                if (boot && m2 == 0u) atomicOr(T.err, kErrEmpty);
                // targets = r + gamma * max * (1 - terminated); delta = lr * (targets - prediction)   (QLO:884-891)
                const float target = __fadd_rn(F.tr_r[ii], boot ? __fmul_rn(F.gamma, m) : 0.0f);
                if (active) F.tr_r[ii] = __fmul_rn(lr, __fsub_rn(target, __ldcg(T.tr_p + ii)));
            }
            grid.sync();
            if (clk && k < 10) F.phase_ns[2 + 3 * k] = global_ns();
            for (int i = tid; i < n; i += nthreads) atomicAdd(T.q + (size_t)cur[i] * T.ld + (F.tr_a[i] & 0x7F), F.tr_r[i]);
            grid.sync();
            if (clk && k < 10) F.phase_ns[3 + 3 * k] = global_ns();
            continue;
        }

        // ---------------- phase B1: exact sequential TD update, first pass (agents that wait for nobody finish)
        // Tiles are claimed from a counter (crowded tiles cost several times more than sparse ones); the next claim
        // is issued before the current tile is processed so that its round trip is hidden.
        {
            int* claim = F.tile_counter + 4 + (k & 1);
            int base = 0;
            if (lane == 0) base = atomicAdd(claim, 32);
            base = __shfl_sync(kFull, base, 0);
            while (base < n) {
                int next_base = 0;
                if (lane == 0) next_base = atomicAdd(claim, 32);
                const int i = base + lane;
                const bool active = i < n;
                const int ii = active ? i : 0;
                const uint8_t at = F.tr_a[ii];
                const int s2 = nxt[ii];
                const uint32_t ew = (ENV == 1) ? F.envw[ii] : 0u;
                const uint32_t m2 = F.use_masks ? env_mask<ENV>(s2, ew, T.A, F.env_seed) : full;
                const int kind = learn_first_pass<LPR>(T, s_best + threadIdx.x, active, i, cur[ii], at & 0x7F, F.tr_r[ii],
                                                       __ldcg(T.tr_p + ii), s2, (at & 0x80) != 0, m2, lr, F.gamma, epoch);
                store_tile_masks(T, base >> 5, kind, F.tile_counter + 2 + (k & 1));
                base = __shfl_sync(kFull, next_base, 0);
            }
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[2 + 3 * k] = global_ns();

        // ---------------- phase B2: deferred agents, in increasing agent order, polling their predecessors
        second_pass_all<LPR>(T, s_best + threadIdx.x, F.tile_counter + (k & 1), __ldcg(F.tile_counter + 2 + (k & 1)) != 0, n, cur, nxt,
                             [&](int i, int s2) {
                                 const uint32_t ew = (ENV == 1) ? __ldcg(F.envw + i) : 0u;
                                 return F.use_masks ? env_mask<ENV>(s2, ew, T.A, F.env_seed) : full;
                             },
                             lr, F.gamma, epoch);
        if (tid == 0) {  // the other cursor / crowded-agent count are idle until the next step
            F.tile_counter[(k + 1) & 1] = 0;
            F.tile_counter[2 + ((k + 1) & 1)] = 0;
            F.tile_counter[4 + ((k + 1) & 1)] = 0;
            T.spill_next[epoch & 1u] = 0;  // the next step hands out slots from the other counter, the one after it from this one
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[3 + 3 * k] = global_ns();
    }
    // leave the current observation in F.st_a
    if (F.steps & 1) {
        for (int i = tid; i < n; i += nthreads) F.st_a[i] = F.st_b[i];
    }
}

}  // namespace qe
