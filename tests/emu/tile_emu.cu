// tests/emu/tile_emu.cu — CPU emulation of the tile-dataflow BandEd kernel (quicked_b200/csrc/qb_tiles.cuh) checked
// against the oracle.  TEST INFRASTRUCTURE: compiled for the HOST only (nvcc -x cu, no device code is run), links
// oracle/libqoracle.so.  The scheduler / tile functions are the same __host__ __device__ code the GPU kernel runs; this
// driver replaces the CTA (slots, packing of tiles onto lanes, the two barriers per round) by plain loops and runs the
// lanes of a round in a shuffled order, which is legal exactly when no tile reads what another tile of the same round
// writes — the property the GPU kernel relies on.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <algorithm>
#include "../../quicked_b200/csrc/qb_tiles.cuh"
#include "../../quicked_b200/csrc/qb_tiletrace.cuh"
#include "../../oracle/quicked_oracle.h"

using namespace qb;

static unsigned long long rng_state = 88172645463325252ull;
static inline unsigned rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (unsigned)(rng_state >> 11); }

static int enc_host(unsigned char c)
{
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}

struct Pair { std::string p, t; };

static Pair gen_pair(int len, double err, int indel_burst)
{
    Pair r;
    r.t.resize(len);
    for (int i = 0; i < len; ++i) r.t[i] = "ACGT"[rnd() & 3];
    r.p = r.t;
    int nerr = (int)(len * err + 0.999);
    for (int e = 0; e < nerr; ++e) {
        if (r.p.empty()) break;
        const int kind = rnd() % 3, pos = rnd() % r.p.size();
        if (kind == 0) r.p[pos] = "ACGT"[rnd() & 3];
        else if (kind == 1) r.p.erase(pos, 1);
        else r.p.insert(r.p.begin() + pos, "ACGT"[rnd() & 3]);
    }
    for (int b = 0; b < indel_burst; ++b) {
        const int L = 20 + rnd() % 200;
        if ((int)r.p.size() > L + 2 && (rnd() & 1)) r.p.erase(rnd() % (r.p.size() - L), L);
        else { std::string ins(L, 'A'); for (auto &c : ins) c = "ACGT"[rnd() & 3]; r.p.insert(rnd() % (r.p.size() + 1), ins); }
    }
    if (r.p.empty()) r.p = "A";
    return r;
}

// Host pools of one emulated launch
struct Emu {
    std::vector<BandTask> tasks;
    std::vector<unsigned char> codes;
    std::vector<unsigned char> raw;
    std::vector<u64> peq;
    std::vector<TileRec> recs;
    std::vector<int2> ranges;
    std::vector<int> scores;
    std::vector<u64> state;
    std::vector<BandOut> outs;
    std::vector<int> punt;
    std::vector<u64> ttext;
    int punt_count = 0;
};

static void build_peq(std::vector<u64> &peq, i64 off, const unsigned char *codes, int m, int rev)
{
    const int nblk = (m + 63) / 64, nbp = nblk + 2;
    for (int blk = 0; blk < nbp; ++blk)
        for (int c = 0; c < kPeqStride; ++c) {
            u64 w = 0;
            if (c < kAlpha && blk < nblk)
                for (int i = 0; i < 64; ++i) {
                    const int row = blk * 64 + i;
                    if (row >= m) w |= 1ull << i;
                    else if ((codes[rev ? m - 1 - row : row] & 7) == c) w |= 1ull << i;
                }
            if (c == kAlpha && blk < nblk)                            // rows holding a character outside "ACGTN"
                for (int i = 0; i < 64; ++i) {
                    const int row = blk * 64 + i;
                    if (row < m && (codes[rev ? m - 1 - row : row] & 8)) w |= 1ull << i;
                }
            peq[off + (i64)blk * kPeqStride + c] = w;
        }
}

template <bool FULL>
static u64 emulate(Emu &E, int RB, int nslots, int L, bool shuffle)
{
    // tile-text pool (device: k_tile_classes + k_tile_text)
    std::vector<u64> ttext;
    for (auto &t : E.tasks) {
        const int ncols = FULL ? t.n : t.finish;
        const int nw = (ncols + 63) / 64 * 8;
        t.tt_off = (i64)ttext.size();
        unsigned odd = 0;
        for (int w = 0; w < nw; ++w) {
            u64 v = 0;
            for (int b = 0; b < 8; ++b) { const int col = 8 * w + b; const unsigned rawc = col < ncols ? E.codes[t.t_off + (t.rev ? t.n - 1 - col : col)] : 4u; odd |= rawc & 8u; v |= (u64)(rawc & 7u) << (8 * b); }
            ttext.push_back(v);
        }
        ttext.push_back(odd ? 1 : 0);
    }
    E.ttext = ttext;
    TilePools P;
    P.ttext = E.ttext.data();
    P.tasks = E.tasks.data(); P.codes = E.codes.data(); P.peq = E.peq.data(); P.recs = E.recs.data();
    P.ranges = E.ranges.data(); P.scores = E.scores.data(); P.state = E.state.data(); P.outs = E.outs.data();
    P.punt_list = E.punt.data(); P.punt_count = &E.punt_count; P.rec_sub = 0;
    std::vector<TileSlot> slots(nslots);
    std::vector<std::vector<unsigned char>> arena(nslots, std::vector<unsigned char>(tile_slot_arena_bytes(RB)));
    std::vector<TileRings> rings(nslots);
    for (int s = 0; s < nslots; ++s) { slots[s].task = -1; slots[s].cnt = 0; slots[s].koff = 0; rings[s] = tile_rings(arena[s].data(), RB); }
    std::vector<char> need_plan(nslots, 0);
    std::vector<u64> eq((size_t)kAlpha * L);
    size_t next = 0;
    u64 ws = 0, rounds = 0, lane_rounds = 0, busy = 0;
    for (;;) {
        // ---- scheduler phase ----
        std::vector<unsigned> plan;
        bool any = false;
        const int rot = (int)(rounds % nslots);
        for (int v = 0; v < nslots; ++v) {
            const int s = (v + rot) % nslots;
            TileSlot &S = slots[s];
            if (S.task >= 0 && S.koff >= S.cnt) { ++S.rho; need_plan[s] = 1; }
            int spin = 0;
            for (;;) {
                if (S.task < 0) {
                    while (next < E.tasks.size()) {
                        const BandTask &tk = E.tasks[next];
                        const BandGeom g = band_geometry(tk.m, tk.n, tk.cutoff);
                        const int B = (int)(FULL ? g.Bc : g.Bs);
                        if (tile_ring_for(B) > RB) { fprintf(stderr, "task %zu needs ring %d > %d\n", next, tile_ring_for(B), RB); exit(2); }
                        tile_slot_load<FULL>(S, rings[s], tk, (int)next, P);
                        need_plan[s] = 1;
                        ++next;
                        break;
                    }
                    if (S.task < 0) break;
                }
                if (need_plan[s]) {
                    if (!S.state) tile_plan_round<FULL>(S, rings[s], P);
                    S.koff = 0;
                    if (S.state == 1) { tile_slot_finish<FULL>(S, rings[s], P); ws += S.ws; S.task = -1; continue; }
                    if (S.state == 2) { E.punt[E.punt_count++] = S.task; S.task = -1; continue; }
                    if (S.cnt == 0) { ++S.rho; if (++spin > 100000) { fprintf(stderr, "slot spins without a due tile\n"); exit(3); } continue; }
                    need_plan[s] = 0;
                }
                any = true;
                {
                    const int want = S.cnt - S.koff, room = L - (int)plan.size();
                    const int take = want < room ? want : room;
                    for (int i = 0; i < take; ++i) plan.push_back(((unsigned)s << 24) | (unsigned)(S.kmin + S.koff + i));
                    S.koff += take;
                }
                break;
            }
        }
        if (!any) break;
        if (plan.empty()) { fprintf(stderr, "deadlock: no tile ready\n"); exit(3); }
        // ---- compute phase: every lane's tile_begin, the mid-pass barrier, every lane's tile_end ----
        if (shuffle) for (size_t i = plan.size(); i > 1; --i) std::swap(plan[i - 1], plan[rnd() % i]);
        std::vector<TileIn> tin(plan.size());
        for (size_t g = 0; g < plan.size(); ++g) {
            const int s = (int)(plan[g] >> 24), k = (int)(plan[g] & 0xffffffu);
            tile_begin<FULL>(slots[s], rings[s], k, P, tin[g]);
        }
        if (shuffle) {      // tile_end in another order than tile_begin
            std::vector<size_t> ord(plan.size());
            for (size_t i = 0; i < ord.size(); ++i) ord[i] = i;
            for (size_t i = ord.size(); i > 1; --i) std::swap(ord[i - 1], ord[rnd() % i]);
            for (size_t g : ord) tile_end<FULL, 0>(slots[(int)(plan[g] >> 24)], rings[(int)(plan[g] >> 24)], tin[g], eq.data() + g, L, P);
        } else
            for (size_t g = 0; g < plan.size(); ++g) tile_end<FULL, 0>(slots[(int)(plan[g] >> 24)], rings[(int)(plan[g] >> 24)], tin[g], eq.data() + g, L, P);
        ++rounds; lane_rounds += L; busy += plan.size();
    }
    fprintf(stderr, "  [emu] %zu tasks, RB %d, %d slots, %d lanes: %llu rounds, lane utilisation %.3f, %d punts\n", E.tasks.size(), RB, nslots, L,
            (unsigned long long)rounds, lane_rounds ? (double)busy / lane_rounds : 0.0, E.punt_count);
    return ws;
}

struct Case { Pair pr; i64 cutoff; int finish; int rev; };

static int check_score_mode(std::vector<Case> &cases, int nslots, int L)
{
    Emu E;
    const size_t nt = cases.size();
    E.tasks.resize(nt); E.outs.resize(nt); E.punt.resize(nt);
    i64 raw = 0, peqw = 0, st = 0, sc = 0;
    int RB = 8;
    for (auto &c : cases) raw += c.pr.p.size() + c.pr.t.size();
    E.codes.assign(raw + 64, 4);
    i64 off = 0;
    for (size_t i = 0; i < nt; ++i) {
        Case &c = cases[i];
        BandTask &t = E.tasks[i];
        memset(&t, 0, sizeof t);
        t.p_off = off; for (char ch : c.pr.p) E.codes[off++] = (unsigned char)enc_host(ch);
        t.t_off = off; for (char ch : c.pr.t) E.codes[off++] = (unsigned char)(enc_host(ch) | ((rnd() & 7) == 0 ? 8 : 0));   // odd flag bits must be ignored
        t.m = (int)c.pr.p.size(); t.n = (int)c.pr.t.size(); t.rev = c.rev; t.finish = c.finish; t.cutoff = c.cutoff;
        t.nbp = (t.m + 63) / 64 + 2; t.peq_off = peqw; peqw += (i64)kPeqStride * t.nbp;
        const BandGeom g = band_geometry(t.m, t.n, t.cutoff);
        t.slot = (int)i; t.state_off = st; st += 2 * g.Bs; t.scores_off = sc; sc += (t.m + 63) / 64 + g.Bs + 2;
        RB = std::max(RB, tile_ring_for(g.Bs));
    }
    E.peq.assign(peqw, 0); E.state.assign(st + 1, 0); E.scores.assign(sc + 1, 0);
    for (size_t i = 0; i < nt; ++i) build_peq(E.peq, E.tasks[i].peq_off, E.codes.data() + E.tasks[i].p_off, E.tasks[i].m, E.tasks[i].rev);
    const u64 ws0 = qo_word_steps_total();
    int bad = 0;
    std::vector<std::vector<uint64_t>> opv(nt), omv(nt);
    std::vector<std::vector<int64_t>> osc(nt);
    std::vector<int64_t> oscore(nt), olo(nt), ohi(nt);
    for (size_t i = 0; i < nt; ++i) {
        Case &c = cases[i];
        const BandTask &t = E.tasks[i];
        const BandGeom g = band_geometry(t.m, t.n, t.cutoff);
        std::string p = c.pr.p, tx = c.pr.t;
        if (c.rev) { std::reverse(p.begin(), p.end()); std::reverse(tx.begin(), tx.end()); }
        opv[i].resize(g.Bs + 1); omv[i].resize(g.Bs + 1); osc[i].resize((t.m + 63) / 64 + g.Bs + 2);
        oscore[i] = qo_banded_score(p.data(), t.m, tx.data(), t.n, t.cutoff, c.finish, opv[i].data(), omv[i].data(), osc[i].data(), &olo[i], &ohi[i]);
    }
    const u64 ws_oracle = qo_word_steps_total() - ws0;
    const u64 ws = emulate<false>(E, RB, nslots, L, true);
    std::vector<char> punted(nt, 0);
    for (int q = 0; q < E.punt_count; ++q) punted[E.punt[q]] = 1;
    u64 ws_punt = 0;
    for (size_t i = 0; i < nt; ++i) {
        const BandTask &t = E.tasks[i];
        const BandGeom g = band_geometry(t.m, t.n, t.cutoff);
        if (punted[i]) {   // not comparable: count the oracle's word-steps of this task so the totals still match
            std::string p = cases[i].pr.p, tx = cases[i].pr.t;
            if (cases[i].rev) { std::reverse(p.begin(), p.end()); std::reverse(tx.begin(), tx.end()); }
            const u64 a = qo_word_steps_total();
            qo_banded_score(p.data(), t.m, tx.data(), t.n, t.cutoff, cases[i].finish, nullptr, nullptr, nullptr, nullptr, nullptr);
            ws_punt += qo_word_steps_total() - a;
            continue;
        }
        const BandOut &o = E.outs[i];
        bool ok = (o.score == oscore[i]) && (o.first == olo[i]) && (o.last == ohi[i]);
        for (int j = 0; j < g.Bs && ok; ++j) {
            const bool act = j >= o.first && j <= o.last && j + o.pos_v >= 0;
            const u64 epv = act ? opv[i][j] : 0, emv = act ? omv[i][j] : 0;
            if (E.state[t.state_off + j] != epv || E.state[t.state_off + g.Bs + j] != emv) ok = false;
        }
        const int nsc = (t.m + 63) / 64 + (int)g.Bs + 2;
        for (int j = 0; j < nsc && ok; ++j) if ((int64_t)E.scores[t.scores_off + j] != osc[i][j]) ok = false;
        if (!ok) {
            ++bad;
            if (bad <= 5) fprintf(stderr, "  MISMATCH score-mode task %zu (m %d n %d cutoff %lld finish %d rev %d): score %d vs %lld, first %d vs %lld, last %d vs %lld\n",
                                  i, t.m, t.n, (long long)t.cutoff, t.finish, t.rev, o.score, (long long)oscore[i], o.first, (long long)olo[i], o.last, (long long)ohi[i]);
        }
    }
    if (ws + ws_punt != ws_oracle) { fprintf(stderr, "  MISMATCH word-steps: emu %llu (+%llu punted) vs oracle %llu\n", (unsigned long long)ws, (unsigned long long)ws_punt, (unsigned long long)ws_oracle); ++bad; }
    return bad;
}

// FULL mode: records against the oracle's stored matrix, then the tile traceback against the oracle's op string.
static int check_full_mode(std::vector<Case> &cases, int nslots, int L, int *n_punt_trace)
{
    Emu E;
    const size_t nt = cases.size();
    E.tasks.resize(nt); E.outs.resize(nt); E.punt.resize(nt);
    i64 raw = 0, peqw = 0, sc = 0, recs = 0, rg = 0;
    int RB = 8;
    for (auto &c : cases) raw += c.pr.p.size() + c.pr.t.size();
    E.codes.assign(raw + 64, 4); E.raw.assign(raw + 64, 0);
    i64 off = 0;
    for (size_t i = 0; i < nt; ++i) {
        Case &c = cases[i];
        BandTask &t = E.tasks[i];
        memset(&t, 0, sizeof t);
        auto stored = [](char ch) { const bool plain = ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'N'; return (unsigned char)(enc_host((unsigned char)ch) | (plain ? 0 : 8)); };
        t.p_off = off; for (char ch : c.pr.p) { E.raw[off] = (unsigned char)ch; E.codes[off++] = stored(ch); }
        t.t_off = off; for (char ch : c.pr.t) { E.raw[off] = (unsigned char)ch; E.codes[off++] = stored(ch); }
        t.m = (int)c.pr.p.size(); t.n = (int)c.pr.t.size(); t.rev = 0; t.finish = t.n; t.cutoff = c.cutoff;
        t.nbp = (t.m + 63) / 64 + 2; t.peq_off = peqw; peqw += (i64)kPeqStride * t.nbp;
        const BandGeom g = band_geometry(t.m, t.n, t.cutoff);
        t.slot = (int)i; t.scores_off = sc; sc += (t.m + 63) / 64 + g.Bc + 2;
        t.mat_off = 2 * recs; recs += (i64)((t.n + 63) / 64) * g.Bc;
        t.range_off = rg; rg += t.n / 64 + 2;
        t.ops_off = 0; t.ops_cap = ((t.m + t.n + 15) / 16) * 16;
        RB = std::max(RB, tile_ring_for(g.Bc));
    }
    E.peq.assign(peqw, 0); E.scores.assign(sc + 1, 0); E.recs.assign(recs + 1, TileRec{0x5555555555555555ull, 0x3333333333333333ull, {1, 2, 3, 4}});
    E.ranges.assign(rg + 1, make_int2(-9, -9));
    for (size_t i = 0; i < nt; ++i) build_peq(E.peq, E.tasks[i].peq_off, E.codes.data() + E.tasks[i].p_off, E.tasks[i].m, 0);
    emulate<true>(E, RB, nslots, L, true);
    std::vector<char> punted(nt, 0);
    for (int q = 0; q < E.punt_count; ++q) punted[E.punt[q]] = 1;
    int bad = 0;
    for (size_t i = 0; i < nt; ++i) {
        if (punted[i]) continue;
        const BandTask &t = E.tasks[i];
        const BandGeom g = band_geometry(t.m, t.n, t.cutoff);
        const int B = (int)g.Bc, n = t.n, m = t.m, prolog = (int)g.prolog;
        std::vector<uint64_t> PV((size_t)B * (n + 1)), MV((size_t)B * (n + 1));
        std::vector<int64_t> org(2 * (n / 64 + 2), -9);
        std::vector<char> oops((size_t)m + n + 1);
        qo_banded_full_dump(cases[i].pr.p.data(), m, cases[i].pr.t.data(), n, t.cutoff, PV.data(), MV.data(), org.data(), oops.data());
        bool ok = true;
        const int nshift = n / 64, K = (n + 63) / 64;
        for (int k = 0; k <= nshift && ok; ++k)
            if (E.ranges[t.range_off + k].x != org[2 * k] || E.ranges[t.range_off + k].y != org[2 * k + 1]) {
                ok = false;
                fprintf(stderr, "  range mismatch task %zu k %d: (%d,%d) vs (%lld,%lld)\n", i, k, E.ranges[t.range_off + k].x, E.ranges[t.range_off + k].y, (long long)org[2 * k], (long long)org[2 * k + 1]);
            }
        const u64 *pq = E.peq.data() + t.peq_off;
        for (int k = 0; k < K && ok; ++k) {
            const int first = (int)org[2 * k], last = (int)org[2 * k + 1];
            const int nc = std::min(64, n - 64 * k);
            for (int j = first; j <= last && ok; ++j) {
                const TileRec &r = E.recs[t.mat_off / 2 + (i64)k * B + j];
                if (r.pv0 != PV[(size_t)(64 * k) * B + j] || r.mv0 != MV[(size_t)(64 * k) * B + j]) { ok = false; fprintf(stderr, "  rec start mismatch task %zu k %d j %d\n", i, k, j); break; }
                u64 pv = r.pv0, mv = r.mv0;
                const int b = j + k - prolog;
                const int ob = (b == (m + 63) / 64 - 1 && (m & 63)) ? (m & 63) - 1 : 63;
                for (int s = 0; s < nc; ++s) {
                    const int code = E.codes[t.t_off + 64 * k + s] & 7;
                    const u32 hp = ((s < 32 ? r.cin.p0 : r.cin.p1) >> (31 - (s & 31))) & 1u, hm = ((s < 32 ? r.cin.m0 : r.cin.m1) >> (31 - (s & 31))) & 1u;
                    u32 a, bq;
                    myers_step_at(b < t.nbp ? pq[(i64)b * kPeqStride + code] : 0ull, pv, mv, hp, hm, ob, a, bq);
                    const int col = 64 * k + s + 1;
                    if (s < 63) { if (pv != PV[(size_t)col * B + j] || mv != MV[(size_t)col * B + j]) { ok = false; fprintf(stderr, "  recompute mismatch task %zu k %d j %d s %d\n", i, k, j, s); break; } }
                    else if (k + 1 <= nshift && j - 1 >= org[2 * (k + 1)]) { if (pv != PV[(size_t)col * B + j - 1] || mv != MV[(size_t)col * B + j - 1]) { ok = false; fprintf(stderr, "  recompute (shifted col) mismatch task %zu k %d j %d\n", i, k, j); break; } }
                }
            }
        }
        // ---- tile traceback ----
        if (ok) {
            std::vector<u32> ops(t.ops_cap / 16 + 2, 0);
            LeafOut lo; memset(&lo, 0, sizeof lo);
            u32 planes[kTraceCols];
            u64 eqs[kAlpha];
            const int rc = tile_traceback(t, E.recs.data() + t.mat_off / 2, E.ranges.data() + t.range_off, E.ttext.data(), E.raw.data(), E.peq.data(),
                                          ops.data(), planes, 1, eqs, 1, lo);
            if (rc != 0) { if (n_punt_trace) ++*n_punt_trace; }
            else {
                std::string got;
                for (int q = t.ops_cap - lo.n_ops; q < t.ops_cap; ++q) got.push_back("MXID"[(ops[q >> 4] >> (2 * (q & 15))) & 3]);
                if (got != std::string(oops.data())) {
                    ok = false;
                    size_t d = 0; while (d < got.size() && d < strlen(oops.data()) && got[got.size() - 1 - d] == oops[strlen(oops.data()) - 1 - d]) ++d;
                    fprintf(stderr, "  traceback mismatch task %zu (m %d n %d cutoff %lld): %zu vs %zu ops, first difference %zu ops from the end\n", i, m, n, (long long)t.cutoff, got.size(), strlen(oops.data()), d);
                } else {
                    int cost = 0; for (char ch : got) cost += ch != 'M';
                    if (cost != lo.cost) { ok = false; fprintf(stderr, "  cost mismatch task %zu: %d vs %d\n", i, lo.cost, cost); }
                    // text length of the run-length string
                    int tl = 0; for (size_t a = 0; a < got.size();) { size_t e = a; while (e < got.size() && got[e] == got[a]) ++e; tl += (int)std::to_string(e - a).size() + 1; a = e; }
                    if (lo.text_len != -1 && tl != lo.text_len) { ok = false; fprintf(stderr, "  text_len mismatch task %zu: %d vs %d\n", i, lo.text_len, tl); }
                }
            }
        }
        if (!ok) { ++bad; if (bad <= 5) fprintf(stderr, "  MISMATCH full-mode task %zu (m %d n %d cutoff %lld B %d)\n", i, m, n, (long long)t.cutoff, B); }
    }
    return bad;
}

int main(int argc, char **argv)
{
    if (argc > 7 && !strcmp(argv[1], "util")) {      // util len err cutoff_frac nslots lanes ntasks: lane utilisation of a homogeneous batch
        const int len = atoi(argv[2]); const double err = atof(argv[3]), cf = atof(argv[4]);
        const int nslots = atoi(argv[5]), L = atoi(argv[6]), nt = atoi(argv[7]);
        std::vector<Case> cs;
        for (int i = 0; i < nt; ++i) { Case c; c.pr = gen_pair(len, err, 0); c.cutoff = (i64)(len * cf); c.rev = 0; c.finish = (int)c.pr.t.size(); cs.push_back(c); }
        return check_score_mode(cs, nslots, L);
    }
    int scale = argc > 1 ? atoi(argv[1]) : 1;
    int bad = 0;
    {   // score mode: many shapes, exact and too-narrow bands, partial passes, reversed passes
        std::vector<Case> cs;
        const int lens[] = {1, 5, 63, 64, 65, 100, 127, 128, 129, 300, 1000, 2500, 6000};
        for (int rep = 0; rep < 6 * scale; ++rep)
            for (int len : lens) {
                const double errs[] = {0.0, 0.05, 0.15, 0.3};
                for (double e : errs) {
                    Case c; c.pr = gen_pair(len, e, (rep % 3 == 2 && len >= 1000) ? 2 : 0);
                    const int m = (int)c.pr.p.size(), n = (int)c.pr.t.size();
                    const int ml = std::max(m, n);
                    const int cuts[] = {0, ml / 50, ml / 10, ml / 4, ml};
                    c.cutoff = cuts[rnd() % 5];
                    c.rev = rnd() & 1;
                    c.finish = (rnd() & 1) ? n : 1 + (int)(rnd() % n);
                    cs.push_back(c);
                }
            }
        // ragged pairs
        for (int rep = 0; rep < 40 * scale; ++rep) {
            Case c; c.pr = gen_pair(50 + rnd() % 900, 0.1, 0);
            c.pr.t = c.pr.t.substr(0, 1 + rnd() % c.pr.t.size());
            if (rnd() & 1) std::swap(c.pr.p, c.pr.t);
            const int m = (int)c.pr.p.size(), n = (int)c.pr.t.size();
            c.cutoff = rnd() % (std::max(m, n) + 1); c.rev = rnd() & 1; c.finish = (rnd() & 1) ? n : 1 + (int)(rnd() % n);
            cs.push_back(c);
        }
        fprintf(stderr, "score mode: %zu tasks\n", cs.size());
        bad += check_score_mode(cs, 6, 48);
        bad += check_score_mode(cs, 32, 256);
        bad += check_score_mode(cs, 1, 700);
    }
    {   // FULL mode
        std::vector<Case> cs;
        const int lens[] = {1, 5, 63, 64, 65, 100, 128, 129, 300, 1000, 2500, 5000};
        for (int rep = 0; rep < 4 * scale; ++rep)
            for (int len : lens) {
                const double errs[] = {0.0, 0.05, 0.15, 0.25};
                for (double e : errs) {
                    Case c; c.pr = gen_pair(len, e, (rep % 3 == 2 && len >= 1000) ? 1 : 0);
                    const int m = (int)c.pr.p.size(), n = (int)c.pr.t.size();
                    const int ml = std::max(m, n);
                    const int cuts[] = {ml / 50, ml / 10, ml / 4, ml / 3, ml};
                    c.cutoff = cuts[rnd() % 5]; c.rev = 0; c.finish = n;
                    if (rep % 2 == 1) {                              // lower case, N and junk: equal codes, different bytes
                        for (auto &ch : c.pr.p) { const unsigned x = rnd() % 40; if (x == 0) ch = (char)tolower(ch); else if (x == 1) ch = 'N'; else if (x == 2) ch = 'n'; else if (x == 3) ch = '*'; }
                        for (auto &ch : c.pr.t) { const unsigned x = rnd() % 40; if (x == 0) ch = (char)tolower(ch); else if (x == 1) ch = 'N'; else if (x == 2) ch = 'R'; }
                    }
                    cs.push_back(c);
                }
            }
        for (int rep = 0; rep < 30 * scale; ++rep) {
            Case c; c.pr = gen_pair(50 + rnd() % 900, 0.1, 0);
            c.pr.t = c.pr.t.substr(0, 1 + rnd() % c.pr.t.size());
            if (rnd() & 1) std::swap(c.pr.p, c.pr.t);
            const int m = (int)c.pr.p.size(), n = (int)c.pr.t.size();
            c.cutoff = rnd() % (std::max(m, n) + 1); c.rev = 0; c.finish = n;
            cs.push_back(c);
        }
        fprintf(stderr, "full mode: %zu tasks\n", cs.size());
        int pt = 0;
        bad += check_full_mode(cs, 8, 64, &pt);
        fprintf(stderr, "  traceback punts (walk left the live band / slice exhausted): %d of %zu\n", pt, cs.size());
        bad += check_full_mode(cs, 32, 300, nullptr);
    }
    if (bad) { fprintf(stderr, "FAILED: %d mismatches\n", bad); return 1; }
    printf("tile emulation OK\n");
    return 0;
}
