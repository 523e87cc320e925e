// pg_travel.cpp -- SURVEY 8a row B9: the min-length path traversal of the A-Bruijn graph and the assembly of the walks
// (`PAssembly::testTravel5`, PAGraph/src/tools/graph/PAssembly.cpp:11-336, over `PAlgorithm::travelSequence`,
// graph/PAlgorithm.cpp:145-426).  Every walk is a sequential pointer chase of a few 10^4 steps and the walks of one
// config block number 2 x contigs, so this stage runs on the host ("replicas only") over the CSR graph that
// ag2_pg_graph_fetch brings back from the device; one host thread per (contig, orientation).
//
// Own data model, not the reference's classes: a position of a vertex is the slot index p of the CSR position arrays
// (the reference's PANode = (vertex, index in the vertex) maps one-to-one to p), a walk is a vector of (vertex, slot,
// step).  All integer widths and the double expressions follow the reference statement by statement because the output
// files must be byte-identical; the comments name the lines.  Paths are relative to PAGraph/src/tools/.
#include "../../include/ag2_b200.h"
#include "../../include/ag2_pagraph.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

namespace {

typedef uint32_t Pos;                            // PABruijnGraph::PosType, graph/PABruijnGraph.hpp:26

struct Mapper {                                  // position/PositionMapper.cpp:8-64
    std::vector<uint64_t> start, sizes;
    Mapper(const int64_t* offs, int64_t n)
    {
        for (int64_t i = 0; i < n; ++i) sizes.push_back((uint64_t)(offs[i + 1] - offs[i]));
        if (n == 0) return;
        start.push_back(sizes[0]);
        for (int64_t i = 1; i < n; ++i) start.push_back(start.back() + 3 * sizes[i - 1] + std::max(sizes[i - 1], sizes[i]));
        start.push_back(start.back() + 4 * sizes[n - 1]);
    }
    uint64_t to_single(int64_t idx, int64_t pos) const          // dualToSingle :40-45
    {
        if (idx == 0) return 0;
        const int64_t i = idx > 0 ? idx - 1 : -idx - 1;
        return start[i] + (idx > 0 ? 0 : 2 * sizes[i]) + (uint64_t)pos;
    }
    std::pair<int64_t, int64_t> to_dual(uint64_t single) const  // singleToDual :47-64; the offset is unsigned there
    {
        if (single == 0) return {0, 0};
        auto it = std::upper_bound(start.begin(), start.end(), single);
        if (it != start.begin()) --it;
        int64_t idx = it - start.begin();
        uint64_t off = single - *it;
        if ((size_t)idx >= sizes.size()) return {idx + 1, (int64_t)off};   // beyond the last sequence: never produced by the build
        if (off >= 2 * sizes[idx]) {
            off -= 2 * sizes[idx];
            idx = -(idx + 1);
        } else {
            ++idx;
        }
        return {idx, (int64_t)off};
    }
    uint64_t size(int64_t idx) const             // :66-70
    {
        if (idx == 0) return 0;
        return sizes[idx > 0 ? idx - 1 : -idx - 1];
    }
};

struct Node {                                    // PABruijnNode: vertex index + position slot
    int64_t v, p;
};
struct Step {
    Node n;
    int dist;
};
typedef std::vector<Step> Walk;                  // PAlgorithm::TravelSequence

enum Grade { Oops, Skip, Good, Excellent, Amazing };          // PABruijnGraph::MatchGrade
enum Status { End, Branch, Limit, Leap };                     // PAlgorithm::NodeStatus

struct PosRange {                                // the (min, max) "ctg pos table", PAlgorithm.cpp:71-101
    Pos lo = std::numeric_limits<Pos>::max(), hi = 0;
    bool has(Pos p) const { return p >= lo && p <= hi; }
    void add(Pos p)
    {
        if (p == 0) return;
        lo = std::min(lo, p);
        hi = std::max(hi, p);
    }
};

inline unsigned base_code(char c)                // seq/CompressedSeq.cpp:16-26
{
    switch (c) {
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 0;
    }
}

struct Seqs {
    const ag2_pg_seqs* s;
    int64_t len(int64_t i) const { return s->offs[i + 1] - s->offs[i]; }
    std::string oriented(int64_t i, bool forward) const      // CompressedSeq::toString :56-73
    {
        const int64_t n = len(i);
        const char* b = s->bases + s->offs[i];
        std::string out((size_t)n, 'A');
        for (int64_t x = 0; x < n; ++x) {
            const unsigned c = base_code(b[x]);
            if (forward) out[x] = "ACGT"[c];
            else out[n - 1 - x] = "TGCA"[c];
        }
        return out;
    }
    char base_at(int64_t i, uint64_t idx, bool forward) const // CompressedSeq::baseAt :75-87
    {
        const uint64_t n = (uint64_t)len(i);
        if (idx >= n) return 'N';
        const char* b = s->bases + s->offs[i];
        return forward ? "ACGT"[base_code(b[idx])] : "TGCA"[base_code(b[n - 1 - idx])];
    }
};

class Traveller {
public:
    const ag2_pg_graph_view& g;
    Seqs ctgs, refs;
    Mapper cmap, rmap;
    int64_t deviation;                           // 2 * epsilon, PGM/pagraph.cpp:251
    double error_rate, start_split;
    uint64_t min_len;
    unsigned threads;

    Traveller(const ag2_pg_graph_view& gv, const ag2_pg_seqs* c, const ag2_pg_seqs* r, const ag2_pg_travel_params& p)
        : g(gv), ctgs{c}, refs{r}, cmap(c->offs, c->n), rmap(r->offs, r->n), deviation(p.deviation), error_rate(p.error_rate),
          start_split(p.start_split), min_len((uint64_t)p.min_len), threads((unsigned)std::max(1, p.threads))
    {
    }

    int64_t vertex_of(uint64_t code) const       // PABruijnGraph::searchDenseIndex :98-104; the codes are sorted (B2), rank = index
    {
        const uint64_t* e = g.codes + g.n_vertices;
        const uint64_t* it = std::lower_bound(g.codes, e, code);
        return it != e && *it == code ? it - g.codes : -1;
    }

    Pos cpos(const Node& n) const { return g.ctg[n.p]; }
    Pos rpos(const Node& n) const { return g.ref[n.p]; }
    unsigned abundance(const Node& n) const { return g.count[n.p]; }

    std::string kmer(int64_t v) const            // KmerHelper::code2Kmer, kmer/KmerHelper.cpp:28-38
    {
        std::string s((size_t)g.k, 'a');
        uint64_t c = g.codes[v];
        for (int i = 0; i < g.k; ++i) {
            s[g.k - 1 - i] = "ACGT"[c & 3];
            c >>= 2;
        }
        return s;
    }

    // ---- PABruijnGraph::isPosSimilar / isEdgeSimilar / checkPosition (graph/PABruijnGraph.cpp:385-400, 143-163) ----
    static void pos_similar(Pos lc, Pos lr, Pos rc, Pos rr, uint64_t dev, bool& s1, bool& s2)
    {
        s1 = lc != 0 && rc != 0 && (uint64_t)(Pos)(std::max(lc, rc) - std::min(lc, rc)) <= dev;
        s2 = lr != 0 && rr != 0 && (uint64_t)(Pos)(std::max(lr, rr) - std::min(lr, rr)) <= dev;
    }
    static void edge_similar(Pos lc, Pos lr, Pos rc, Pos rr, int dist, uint64_t dev, double er, bool& s1, bool& s2)
    {
        const Pos tc = lc != 0 ? (Pos)(lc + (Pos)dist) : 0, tr = lr != 0 ? (Pos)(lr + (Pos)dist) : 0;
        pos_similar(tc, tr, rc, rr, dev, s1, s2);
        s1 = s1 || (lc != 0 && rc != 0 && std::abs(1.0 - ((Pos)(rc - lc) * 1.0 / dist)) <= er);
        s2 = s2 || (lr != 0 && rr != 0 && std::abs(1.0 - ((Pos)(rr - lr) * 1.0 / dist)) <= er);
    }
    static Grade grade(Pos c1, Pos r1, Pos c2, Pos r2, Pos dist, Pos dev, double er)
    {
        bool s1, s2;
        edge_similar(c1, r1, c2, r2, (int)dist, dev, er, s1, s2);
        s1 = s1 || std::abs(1.0 - ((Pos)(c2 - c1) * 1.0 / dist)) <= er;
        s2 = s2 || std::abs(1.0 - ((Pos)(r2 - r1) * 1.0 / dist)) <= er;
        if (c1 == 0 || c2 == 0) return s2 ? (c2 != 0 ? Excellent : (c1 != 0 ? Skip : Good)) : Oops;
        if (r1 == 0 || r2 == 0) return s1 ? (r2 != 0 ? Excellent : Good) : Oops;
        return (s1 && s2) ? Amazing : (s1 ? Excellent : (s2 ? Skip : Oops));
    }
    bool edge_similar_ctg(const Node& a, const Step& b) const
    {
        bool s1, s2;
        edge_similar(cpos(a), rpos(a), cpos(b.n), rpos(b.n), b.dist, (uint64_t)deviation, error_rate, s1, s2);
        return s1;
    }

    // PABruijnGraph::searchSuccessors :165-197: every position of every child vertex that is not Oops, in edge order
    void successors(const Node& parent, std::vector<Step>& out) const
    {
        const Pos c1 = cpos(parent), r1 = rpos(parent);
        for (int64_t e = g.edge_off[parent.v]; e < g.edge_off[parent.v + 1]; ++e) {
            const int64_t to = g.edge_to[e];
            const int step = g.edge_step[e];
            for (int64_t p = g.pos_off[to]; p < g.pos_off[to + 1]; ++p)
                if (grade(c1, r1, g.ctg[p], g.ref[p], (Pos)step, (Pos)deviation, error_rate) != Oops) out.push_back({{to, p}, step});
        }
    }

    struct Range {
        int64_t lo, hi;
        bool outside(Pos p) const { return p != 0 && ((int64_t)p < lo || (int64_t)p >= hi); }
    };

    // PAlgorithm::classifySuccessors, graph/PAlgorithm.tcc:37-90
    template <typename Filter>
    void classify(std::vector<Step>& results, const Node& node, Range range, bool can_leap, double leap_min, Filter keep) const
    {
        std::vector<Step> all, cand;
        successors(node, all);
        for (const Step& s : all)
            if (keep(node, s)) cand.push_back(s);
        std::vector<size_t> cls[4];              // amazing, excellent, good, skip
        const Pos c1 = cpos(node), r1 = rpos(node);
        for (size_t i = 0; i < cand.size(); ++i) {
            const Pos c2 = cpos(cand[i].n), r2 = rpos(cand[i].n);
            const Grade gr = grade(c1, r1, c2, r2, (Pos)cand[i].dist, (Pos)deviation, error_rate);
            const bool leap = range.outside(c2);
            if (leap) {
                const auto dual = cmap.to_dual(c2);
                if ((double)dual.second > (double)cmap.size(dual.first) * leap_min) continue;
            }
            if (!can_leap && leap) continue;
            if (gr == Amazing || leap) cls[0].push_back(i);
            else if (gr == Excellent) cls[1].push_back(i);
            else if (gr == Good) cls[2].push_back(i);
            else if (can_leap && gr == Skip) cls[3].push_back(i);
        }
        const std::vector<size_t>& chosen = !cls[0].empty() ? cls[0] : (!cls[1].empty() ? cls[1] : (!cls[2].empty() ? cls[2] : cls[3]));
        for (size_t i : chosen) results.push_back(cand[i]);
    }

    // PAlgorithm::walkStraight, PAlgorithm.tcc:94-170 (limitation = 0)
    template <typename Filter>
    Status walk_straight(const Step& first, Walk& path, Range range, uint64_t has_size, uint64_t split_size, double split_min,
                         Filter parent_keep) const
    {
        std::unordered_set<int64_t> seen;
        PosRange local;
        uint64_t now = (uint64_t)(int64_t)first.dist;
        path.push_back(first);
        if (range.outside(cpos(first.n))) return Leap;
        local.add(cpos(path.front().n));
        seen.insert(first.n.p);
        auto keep = [&](const Node& from, const Step& s) {
            return parent_keep(from, s) && seen.count(s.n.p) == 0 &&
                   (cpos(s.n) == 0 || edge_similar_ctg(from, s) || !local.has(cpos(s.n)));
        };
        std::vector<Step> next;
        for (;;) {
            next.clear();
            classify(next, path.back().n, range, has_size + now >= split_size, split_min, keep);
            if (next.empty()) return End;
            if (next.size() > 1) return Branch;
            const Step s = next.front();
            seen.insert(s.n.p);
            local.add(cpos(s.n));
            path.push_back(s);
            now += (uint64_t)(int64_t)s.dist;
            if (range.outside(cpos(s.n))) return Leap;
        }
    }

    // PAlgorithm::graphTravel, PAlgorithm.tcc:173-298
    template <typename Filter>
    Walk graph_travel(const Node& start, Range range, uint64_t has_size, uint64_t split_size, double split_min, Filter parent_keep) const
    {
        PosRange travelled;
        std::unordered_set<int64_t> seen;
        Walk seq;
        uint64_t now = (uint64_t)g.k;
        Walk path;
        std::vector<Walk> paths;
        travelled.add(cpos(start));
        auto keep = [&](const Node& from, const Step& s) {
            return parent_keep(from, s) && seen.count(s.n.p) == 0 &&
                   (cpos(s.n) == 0 || edge_similar_ctg(from, s) || !travelled.has(cpos(s.n)));
        };
        walk_straight(Step{start, g.k}, path, range, has_size + now, split_size, split_min, keep);
        paths.push_back(path);
        size_t chosen_path = 0;
        std::vector<Step> next;
        for (;;) {
            const Walk& cp = paths[chosen_path];
            for (const Step& s : cp) {
                seq.push_back(s);
                seen.insert(s.n.p);
                now += (uint64_t)(int64_t)s.dist;
            }
            for (const Step& s : cp) travelled.add(cpos(s.n));
            const Node last = seq.back().n;
            if (range.outside(cpos(last))) break;
            next.clear();
            classify(next, last, range, has_size + now >= split_size, split_min, keep);
            std::vector<std::pair<size_t, size_t>> leap, branch, tips;
            paths.clear();
            for (size_t i = 0; i < next.size(); ++i) {
                path.clear();
                const Status st = walk_straight(next[i], path, range, has_size + now, split_size, split_min, keep);
                paths.push_back(path);
                if (st == Leap) leap.emplace_back(i, path.size());
                else if (st == End) tips.emplace_back(i, path.size());
                else branch.emplace_back(i, path.size());
            }
            if (leap.empty() && tips.empty() && branch.empty()) break;
            if (!leap.empty()) {
                chosen_path = leap.front().first;
            } else if (!branch.empty()) {        // highest abundance, first on ties
                size_t c = 0;
                for (size_t i = 1; i < branch.size(); ++i)
                    if (abundance(next[branch[i].first].n) > abundance(next[branch[c].first].n)) c = i;
                chosen_path = branch[c].first;
            } else {                             // longest tip, first on ties
                size_t c = 0;
                for (size_t i = 1; i < tips.size(); ++i)
                    if (tips[i].second > tips[c].second) c = i;
                chosen_path = tips[c].first;
            }
        }
        return seq;
    }

    static uint64_t walk_size(const Walk& w)     // PAlgorithm::seqSize :491-497
    {
        uint64_t n = 0;
        for (const Step& s : w) n += (uint64_t)(int64_t)s.dist;
        return n;
    }

    // PAlgorithm::appendSeq :111-142
    int64_t append_walk(Walk& base, const Walk& tail) const
    {
        if (tail.empty()) return 0;
        int64_t dlen = 0;
        const Step& head = tail.front();
        int dist = g.k;
        while (!base.empty() && (cpos(base.back().n) == 0 || cpos(head.n) <= cpos(base.back().n))) {
            dlen -= base.back().dist;
            base.pop_back();
        }
        if (!base.empty()) dist = (int)(Pos)(cpos(head.n) - cpos(base.back().n));
        for (const Step& s : tail) {
            dlen += s.dist;
            base.push_back(s);
        }
        Step& joint = base[base.size() - tail.size()];
        dlen -= joint.dist - dist;
        joint.dist = dist;
        return dlen;
    }

    static uint64_t edit_distance(const std::string& a, const std::string& b)   // PAlgorithm::editDistance :46-69
    {
        std::vector<uint64_t> prev(b.size() + 1), cur(b.size() + 1);
        for (size_t j = 0; j <= b.size(); ++j) prev[j] = j;
        for (size_t i = 1; i <= a.size(); ++i) {
            cur[0] = i;
            for (size_t j = 1; j <= b.size(); ++j) {
                cur[j] = std::min(prev[j] + 1, cur[j - 1] + 1);
                cur[j] = std::min(cur[j], prev[j - 1] + (a[i - 1] == b[j - 1] ? 0 : 1));
            }
            prev.swap(cur);
        }
        return prev[b.size()];
    }

    // PAlgorithm::filterSequence :28-44
    static void filter_tail(Walk& seq, const Traveller& t)
    {
        const size_t window = 10;
        if (seq.size() < window) return;
        for (size_t i = seq.size() - seq.size() / 90; i < seq.size() - window + 1; ++i) {
            const Pos a = t.cpos(seq[i].n), b = t.cpos(seq[std::min(seq.size(), i + window) - 1].n);
            if (b != 0 && a != 0 && b < a) {
                seq.resize(i + 1);
                return;
            }
        }
    }

    // the start vertices on the contig: PAlgorithm::searchPANode / searchPANode2, PAlgorithm.tcc:300-365
    template <typename Accept>
    void search_starts(const std::vector<std::pair<int64_t, uint64_t>>& on_ctg, std::vector<Node>& out, bool only_first, bool windowed,
                       uint64_t pos, Accept accept) const
    {
        std::unordered_set<int64_t> taken;
        const uint64_t left = pos - std::min(pos, 1000 * (uint64_t)deviation), right = pos + 1000 * (uint64_t)deviation;
        for (const auto& a : on_ctg) {
            if (windowed) {
                if (a.second < left) continue;
                if (a.second > right) break;
            }
            for (int64_t p = g.pos_off[a.first]; p < g.pos_off[a.first + 1]; ++p) {
                const auto d1 = cmap.to_dual(g.ctg[p]);
                if (taken.count(p) == 0 && accept(a.second, d1.first, (uint64_t)d1.second)) {
                    out.push_back({a.first, p});
                    taken.insert(p);
                }
            }
            if (!out.empty() && only_first) break;
        }
    }

    // PAlgorithm::travelSequence :145-426
    Walk travel(int64_t ctg, bool forward) const
    {
        const size_t top_k = std::min(threads, 8U);
        std::unordered_set<int64_t> done;
        PosRange done_range;
        const int64_t chosen = forward ? ctg + 1 : -ctg - 1;
        const std::string text = ctgs.oriented(ctg, forward);
        const uint64_t ctg_len = (uint64_t)ctgs.len(ctg);

        std::vector<std::pair<int64_t, uint64_t>> on_ctg;   // PABruijnGraph::findAll :329-343 over KmerHelper::kmer2Code
        if (text.size() >= (size_t)g.k) {
            const uint64_t mask = (1UL << (g.k * 2)) - 1;
            uint64_t code = 0;
            for (size_t i = 0; i < text.size(); ++i) {
                code = (code << 2) | base_code(text[i]);
                if (i + 1 < (size_t)g.k) continue;
                if (i + 1 > (size_t)g.k) code &= mask;
                const int64_t v = vertex_of(code);
                if (v >= 0) on_ctg.emplace_back(v, (uint64_t)(i + 1 - g.k));
            }
        }

        const uint64_t split_len = (uint64_t)(ctg_len * start_split);
        const double split_min = 1 - start_split;
        const Range range{(int64_t)(Pos)cmap.to_single(chosen, 0), (int64_t)(Pos)cmap.to_single(chosen, (int64_t)ctg_len)};
        const Pos rev_lo = (Pos)cmap.to_single(-chosen, 0), rev_hi = (Pos)cmap.to_single(-chosen, (int64_t)ctg_len);

        auto keep = [&](const Node& from, const Step& s) {
            const Pos c = cpos(s.n);
            return done.count(s.n.p) == 0 && (c == 0 || edge_similar_ctg(from, s) || !done_range.has(c)) &&
                   (c == 0 || (c < rev_lo || c >= rev_hi));
        };

        std::vector<Node> starts;
        search_starts(on_ctg, starts, true, false, 0, [&](uint64_t at, int64_t ci, uint64_t cp) {
            return ci == chosen && std::max(cp, at) - std::min(cp, at) <= (uint64_t)deviation;
        });
        starts.resize(std::min(starts.size(), top_k));

        Walk whole, longest;
        int64_t var_len = 0;
        std::deque<Pos> ctg_q, ref_q;
        const size_t max_q = 4;
        bool final_leap = false;

        while (!starts.empty()) {
            longest.clear();
            uint64_t max_len = 0, choose_ctg = 0, choose_ref = 0;
            bool leap = false;
            std::vector<Walk> walks(starts.size());
            {                                    // MultiThreadTools::multiTraversal over the start vertices
                const unsigned nt = std::max(1U, std::min<unsigned>(threads, (unsigned)starts.size()));
                auto work = [&](unsigned t) {
                    for (size_t i = t; i < starts.size(); i += nt)
                        walks[i] = graph_travel(starts[i], range, (uint64_t)var_len, split_len, split_min, keep);
                };
                if (nt == 1) {
                    work(0);
                } else {
                    std::vector<std::thread> pool;
                    for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work, t);
                    for (auto& th : pool) th.join();
                }
            }
            for (size_t i = 0; i < starts.size(); ++i) {
                const Walk& w = walks[i];
                const uint64_t len = walk_size(w);
                leap = cpos(w.back().n) != 0 && cmap.to_dual(cpos(w.back().n)).first != chosen;
                if (!leap && i > 0 && min_len > 0 && len < min_len) continue;
                if (len > max_len || leap) {
                    max_len = len;
                    longest = w;
                    choose_ctg = (uint64_t)cmap.to_dual(cpos(starts[i])).second;
                    choose_ref = (uint64_t)rmap.to_dual(rpos(starts[i])).second;
                    if (leap) break;
                }
            }
            var_len += append_walk(whole, longest);
            if (choose_ctg != 0) {
                ctg_q.push_back((Pos)choose_ctg);
                while (ctg_q.size() > max_q) ctg_q.pop_front();
            }
            if (choose_ref != 0) {
                ref_q.push_back((Pos)choose_ref);
                while (ref_q.size() > max_q) ref_q.pop_front();
            }
            for (const Step& s : longest) done.insert(s.n.p);
            for (const Step& s : longest) done_range.add(cpos(s.n));

            bool ctg_repeat = false, ref_repeat = false;
            if (ctg_q.size() >= max_q) {
                auto mm = std::minmax_element(ctg_q.begin(), ctg_q.end());
                ctg_repeat = (uint64_t)(Pos)(*mm.second - *mm.first) <= 2 * (uint64_t)deviation;
            }
            if (ref_q.size() >= max_q) {
                auto mm = std::minmax_element(ref_q.begin(), ref_q.end());
                ref_repeat = (uint64_t)(Pos)(*mm.second - *mm.first) <= 2 * (uint64_t)deviation;
            }
            if (ctg_repeat || ref_repeat || leap) {
                if (leap) final_leap = true;
                break;
            }

            uint64_t last_ctg_pos = 0;
            std::string last_kmer;
            bool f1 = false, f2 = false;
            for (auto it = whole.rbegin(); (!f1 || !f2) && it != whole.rend(); ++it) {
                if (!f1 && cpos(it->n) != 0) {
                    const auto d = cmap.to_dual(cpos(it->n));
                    if (d.first == chosen && d.second >= 0) {
                        last_ctg_pos = (uint64_t)d.second;
                        last_kmer = kmer(it->n.v);
                        f1 = true;
                    }
                }
                if (!f2 && rpos(it->n) != 0) f2 = true;
            }
            starts.clear();
            search_starts(on_ctg, starts, false, true, last_ctg_pos, [&](uint64_t, int64_t ci, uint64_t cp) {
                return ci == chosen && std::max(cp, last_ctg_pos) - std::min(cp, last_ctg_pos) <= (uint64_t)deviation;
            });
            {
                size_t n = 0;
                for (size_t i = 0; i < starts.size(); ++i)
                    if (done.count(starts[i].p) == 0) starts[n++] = starts[i];
                starts.resize(n);
            }
            // std::sort by edit distance to the last contig k-mer (:396-400): the same call on precomputed keys makes
            // the same comparisons, so equal keys come out in the order libstdc++ leaves them in the reference
            std::vector<std::pair<uint64_t, Node>> keyed;
            for (const Node& s : starts) keyed.emplace_back(edit_distance(last_kmer, kmer(s.v)), s);
            std::sort(keyed.begin(), keyed.end(),
                      [](const std::pair<uint64_t, Node>& a, const std::pair<uint64_t, Node>& b) { return a.first < b.first; });
            for (size_t i = 0; i < starts.size(); ++i) starts[i] = keyed[i].second;
            starts.resize(std::min(starts.size(), top_k));
        }

        if (!final_leap) filter_tail(whole, *this);
        if (final_leap) {
            const auto d = cmap.to_dual(cpos(whole.back().n));
            if ((uint64_t)std::abs(d.first) == (uint64_t)ctg + 1 ||
                (double)d.second >= (double)(uint64_t)ctgs.len(std::abs(d.first) - 1) * (1 - start_split))
                whole.pop_back();
        }
        return whole;
    }

    // PAlgorithm::seqToString :428-489
    std::string walk_string(const Walk& seq) const
    {
        if (seq.empty()) return "";
        std::string str = kmer(seq[0].n.v);
        const int k = g.k;
        for (size_t i = 1; i < seq.size(); ++i) {
            const Node& a = seq[i - 1].n;
            const Step& b = seq[i];
            bool s1, s2;
            edge_similar(cpos(a), rpos(a), cpos(b.n), rpos(b.n), b.dist, (uint64_t)deviation, error_rate, s1, s2);
            bool use_ctg = s1;
            if (!s1 && !s2) {
                bool p1, p2;
                pos_similar(cpos(a), rpos(a), cpos(b.n), rpos(b.n), (uint64_t)deviation, p1, p2);
                use_ctg = p1;
            }
            const Seqs& db = use_ctg ? ctgs : refs;
            const Mapper& m = use_ctg ? cmap : rmap;
            const auto from = m.to_dual(use_ctg ? cpos(a) : rpos(a)), to = m.to_dual(use_ctg ? cpos(b.n) : rpos(b.n));
            const int kd = b.dist;
            const int64_t pd = to.second - from.second;
            const int64_t sel = std::abs(to.first) - 1;
            const bool fwd = to.first > 0;
            const double move = pd * 1.0 / kd;
            double at = (double)(from.second + k);
            const std::string km = kmer(b.n.v);
            for (int j = 0; j < kd; ++j) {
                if (k - kd + j >= 0) {
                    str.push_back(km[k - kd + j]);
                } else {
                    const double r = std::round(at);
                    const uint64_t rp = r < 9223372036854775808.0 ? (uint64_t)(int64_t)r : (uint64_t)r;
                    const char c = sel >= 0 && sel < db.s->n ? db.base_at(sel, rp, fwd) : 'N';
                    str.push_back((char)std::tolower(c));
                }
                at += move;
            }
        }
        return str;
    }
};

// PAssembly::combatSeq, graph/PAssembly.tcc:3-31: follow the leaps from walk to walk
template <typename F>
void chain(const std::vector<Walk>& walks, const Traveller& t, size_t start, bool forward, F visit)
{
    size_t next = start * 2 + (forward ? 0 : 1);
    uint64_t next_pos = 0;
    std::set<size_t> seen{next};
    for (;;) {
        if (!visit(next / 2, next % 2 == 0, next_pos)) break;
        if (walks[next].empty() || t.cpos(walks[next].back().n) == 0) break;
        const auto d = t.cmap.to_dual(t.cpos(walks[next].back().n));
        next_pos = (uint64_t)d.second;
        next = (size_t)(std::abs(d.first) - 1) * 2 + (d.first > 0 ? 0 : 1);
        if (seen.count(next)) break;
        seen.insert(next);
    }
}

struct Roots {                                   // graph/UnionSet.cpp
    std::vector<size_t> parent;
    explicit Roots(size_t n) : parent(n)
    {
        for (size_t i = 0; i < n; ++i) parent[i] = i;
    }
    size_t find(size_t x) { return x == parent[x] ? x : (parent[x] = find(parent[x])); }
    void join(size_t left, size_t right) { parent[find(right)] = find(left); }
};

}  // namespace

extern "C" {

void ag2_pg_travel_params_default(ag2_pg_travel_params* p)
{
    if (!p) return;
    p->deviation = 20;        // 2 * --epsilon
    p->error_rate = 0.15;
    p->start_split = 0.90;
    p->min_len = 50;
    p->threads = 16;
}

// PAssembly::testTravel5 (graph/PAssembly.cpp:11-336)
int ag2_pg_travel(const ag2_pg_graph_view* gv, const ag2_pg_seqs* ctgs, const ag2_pg_seqs* refs, const int32_t* use_ctg,
                  const uint8_t* use_forward, int64_t n_use, const ag2_pg_travel_params* prm, const char* out_dir, const char* prefix,
                  int32_t* ok_ctg, int64_t* n_ok)
{
    if (!gv || !ctgs || !refs || !prm || !out_dir || !prefix || (n_use > 0 && (!use_ctg || !use_forward)) || gv->k < 1 || gv->k > 31)
        return AG2_EINVAL;
    if (n_ok) *n_ok = 0;
    for (int64_t i = 0; i < n_use; ++i)
        if (use_ctg[i] < 0 || use_ctg[i] >= ctgs->n) return AG2_EINVAL;
    const Traveller T(*gv, ctgs, refs, *prm);
    const std::string dir = std::string(out_dir) + "/" + prefix;

    // std::set<std::pair<std::string, bool>> usedCtg (PGM/pagraph.cpp:209-214): ordered by (name, flag)
    std::set<std::pair<std::string, bool>> used;
    std::map<std::pair<std::string, bool>, int32_t> id_of;
    for (int64_t i = 0; i < n_use; ++i) {
        std::pair<std::string, bool> key(ctgs->names[use_ctg[i]], use_forward[i] != 0);
        used.insert(key);
        id_of[key] = use_ctg[i];
    }
    std::vector<std::pair<std::string, bool>> list(used.begin(), used.end());
    std::vector<Walk> walks((size_t)ctgs->n * 2);
    std::vector<uint64_t> in_deg((size_t)ctgs->n * 2, 0);
    std::vector<int> io_fail(list.size(), 0);

    auto slot = [&](const std::pair<std::string, bool>& key) { return (size_t)id_of[key] * 2 + (key.second ? 0 : 1); };

    // walks are independent; threadNum / 8 of them at a time in the reference (:31), all of them here
    {
        const unsigned nt = std::max(1U, std::min<unsigned>((unsigned)prm->threads, (unsigned)list.size()));
        auto work = [&](unsigned t) {
            for (size_t ii = t; ii < list.size(); ii += nt) {
                const int64_t ci = id_of.find(list[ii])->second;
                const int o = list[ii].second ? 0 : 1;
                Walk& w = walks[(size_t)ci * 2 + o];
                w = T.travel(ci, o == 0);
                FILE* f = fopen((dir + std::to_string(ci) + "_" + std::to_string(o) + ".txt").c_str(), "w");
                if (!f) {
                    io_fail[ii] = 1;
                    continue;
                }
                fprintf(f, "%s\t%llu\n", list[ii].first.c_str(), (unsigned long long)T.ctgs.len(ci));
                for (const Step& s : w) {
                    const auto d1 = T.cmap.to_dual(T.cpos(s.n)), d2 = T.rmap.to_dual(T.rpos(s.n));
                    fprintf(f, "%s,%u,%u,%u\t%d\t%lld,%lld\t%lld,%lld\n", T.kmer(s.n.v).c_str(), T.cpos(s.n), T.rpos(s.n), T.abundance(s.n),
                            s.dist, (long long)d1.first, (long long)d1.second, (long long)d2.first, (long long)d2.second);
                }
                fclose(f);
                if ((double)Traveller::walk_size(w) < (double)(uint64_t)T.ctgs.len(ci) * prm->start_split * 0.9) w.clear();
            }
        };
        if (nt == 1) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work, t);
            for (auto& th : pool) th.join();
        }
    }
    for (int x : io_fail)
        if (x) return AG2_EINVAL;

    auto leap_target = [&](size_t i, size_t& target) {   // the walk a non-empty walk ends in, if it is another one
        if (walks[i].empty()) return false;
        const Pos last = T.cpos(walks[i].back().n);
        if (last == 0) return false;
        const auto d = T.cmap.to_dual(last);
        target = (size_t)(std::abs(d.first) - 1) * 2 + (d.first > 0 ? 0 : 1);
        return target != i;
    };
    for (const auto& key : list) {               // :63-77
        size_t tgt;
        if (leap_target(slot(key), tgt)) ++in_deg[tgt];
    }
    for (const auto& key : used) {               // :119-139
        size_t tgt;
        const size_t i = slot(key);
        if (leap_target(i, tgt) && walks[tgt].empty()) {
            walks[i].pop_back();
            --in_deg[tgt];
        }
    }

    // :150-217 union of chained contigs, one start per component (the longest chain)
    std::map<std::pair<std::string, bool>, size_t> helper;
    std::vector<std::pair<std::string, bool>> table;
    for (const auto& key : used) {
        helper[key] = table.size();
        table.push_back(key);
    }
    std::vector<bool> touched(helper.size(), false);
    Roots roots(helper.size());
    for (const auto& key : used) {
        const size_t i = slot(key);
        if (in_deg[i] > 0 || walks[i].empty()) continue;
        const size_t main_idx = helper[key];
        const size_t ctg_idx = i / 2;
        touched[main_idx] = true;
        chain(walks, T, i / 2, i % 2 == 0, [&](size_t cid, bool fwd, uint64_t) {
            if (cid != ctg_idx || fwd != key.second) {
                const size_t h = helper[{std::string(ctgs->names[cid]), fwd}];   // operator[]: a contig outside the set gets slot 0 there too
                roots.join(h, main_idx);
                if (touched[h]) return false;
                touched[h] = true;
                return true;
            }
            return true;
        });
    }
    // helper may have grown by operator[] on names outside the used set (value 0): the reference then indexes touched /
    // the union set with 0, which is what the lookups above did
    std::vector<std::vector<size_t>> merged(table.size());
    for (size_t i = 0; i < table.size(); ++i) merged[roots.find(i)].push_back(i);
    std::set<std::pair<std::string, bool>> start_set;
    for (const auto& comp : merged) {
        if (comp.empty()) continue;
        uint64_t best = 0;
        size_t chosen = comp.front();
        for (size_t idx : comp) {
            const size_t i = slot(table[idx]);
            if (in_deg[i] > 0 || walks[i].empty()) continue;
            uint64_t len = 0;
            chain(walks, T, i / 2, i % 2 == 0, [&](size_t cid, bool, uint64_t) {
                len += (uint64_t)T.ctgs.len((int64_t)cid);
                return true;
            });
            if (len > best) {
                best = len;
                chosen = idx;
            }
        }
        start_set.insert(table[chosen]);
    }

    // :229-333 one FASTA per start whose chain connects contigs or extends its own
    std::set<std::pair<std::string, bool>> success;
    size_t name_cnt = 0;
    for (const auto& key : start_set) {
        const size_t i = slot(key);
        if (in_deg[i] > 0 || walks[i].empty()) continue;
        const std::string name = std::string(prefix) + std::to_string(name_cnt++);
        std::set<std::pair<size_t, bool>> connected;
        uint64_t max_len = 0, total = 0;
        chain(walks, T, i / 2, i % 2 == 0, [&](size_t cid, bool fwd, uint64_t) {
            connected.emplace(cid, fwd);
            max_len = std::max(max_len, (uint64_t)T.ctgs.len((int64_t)cid));
            total += Traveller::walk_size(walks[cid * 2 + (fwd ? 0 : 1)]);
            return true;
        });
        const bool is_connected = connected.size() > 1 && (double)total > (double)max_len * 1.05;
        const bool is_extended = connected.size() == 1 && (double)Traveller::walk_size(walks[i]) > (double)(uint64_t)T.ctgs.len((int64_t)(i / 2)) * 1.2;
        if (!is_connected && !is_extended) continue;
        const std::string stem = dir + std::to_string(i / 2) + "_" + std::to_string(i % 2);
        FILE* help = fopen((stem + ".help").c_str(), "w");
        FILE* fa = fopen((stem + ".fasta").c_str(), "w");
        FILE* con = fopen((stem + ".con").c_str(), "w");
        if (!help || !fa || !con) {
            if (help) fclose(help);
            if (fa) fclose(fa);
            if (con) fclose(con);
            return AG2_EINVAL;
        }
        fprintf(help, "%llu\n%llu\n", (unsigned long long)total, (unsigned long long)max_len);
        fclose(help);
        fprintf(fa, ">%s\n", name.c_str());
        size_t col = 0;
        uint64_t cmb = 0;
        std::vector<std::pair<std::pair<std::string, bool>, uint64_t>> parts;
        chain(walks, T, i / 2, i % 2 == 0, [&](size_t cid, bool fwd, uint64_t) {
            parts.push_back({{std::string(ctgs->names[cid]), fwd}, (uint64_t)T.ctgs.len((int64_t)cid)});
            for (char ch : T.walk_string(walks[cid * 2 + (fwd ? 0 : 1)])) {
                fputc(ch, fa);
                ++cmb;
                if (++col % 70 == 0) {
                    fputc('\n', fa);
                    col = 0;
                }
            }
            return true;
        });
        if (col > 0) fputc('\n', fa);
        fclose(fa);
        fprintf(con, "%s\t%llu\n", name.c_str(), (unsigned long long)cmb);
        for (const auto& p : parts) fprintf(con, "%s\t%s\t%llu\n", p.first.first.c_str(), p.first.second ? "FORWARD" : "REV", (unsigned long long)p.second);
        fclose(con);
        for (const auto& c : connected) success.emplace(std::string(ctgs->names[c.first]), c.second);
    }
    if (ok_ctg && n_ok) {
        int64_t n = 0;
        for (const auto& s : success) {          // in the order of the reference's std::set
            for (int64_t c = 0; c < ctgs->n; ++c)
                if (s.first == ctgs->names[c]) {
                    ok_ctg[n++] = (int32_t)c;
                    break;
                }
        }
        *n_ok = n;
    }
    return AG2_OK;
}

}  // extern "C"
