/*
 * ag2_pagraph.cpp -- CPU restatement of PAGraph's A-Bruijn graph build.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may call this; nothing under aligngraph2_b200/ does.  It follows the reference statement by statement with plain
 * vectors (one thread, i.e. `pagraph -t 1`, the only deterministic setting of the reference) and is pinned
 * against oracle/_ref/pagraph_dump (the unmodified reference sources + a dump of the dense table) by
 * tests/test_oracle_pagraph.py and the committed fixture tests/golden/pagraph_small.tar.xz.
 *
 * Paths below are relative to /root/reference/PAGraph/src/tools/ (PGM = ../main).
 *
 *   ag2o_pagraph_dump(kmer.bin, ctg.fasta, ref.fasta, pre dir, ctg-to-ref aln, epsilon, cov, out.txt)
 *
 * writes the same text as pagraph_dump:  "#config n ref" then per non-empty vertex
 *   V <dense idx> <code> P <n> {ctg,ref,count}.. E <m> {to,step}..
 */
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

typedef std::pair<int64_t, int64_t> RefPos;                 /* Aligner::ref_pos_t, align/Aligner.hpp:30 */
typedef std::pair<uint32_t, uint32_t> DualPos;              /* PABruijnGraph::DualPos, graph/PABruijnGraph.hpp:27 */

struct SeqDB {                                              /* seq/AutoSeqDatabase.cpp:9-22 */
    std::vector<std::string> name, seq;
    std::unordered_map<std::string, size_t> id;
    bool has(const std::string& n) const { return id.count(n) > 0; }
};

/* seq/CompressedSeq.cpp:7-40 stores 2 bits per base (anything but CcGgTt is A), :57-75 decodes to upper case */
char norm_base(char c)
{
    switch (c) {
    case 'C': case 'c': return 'C';
    case 'G': case 'g': return 'G';
    case 'T': case 't': return 'T';
    default: return 'A';
    }
}

void add_seq(SeqDB& db, const std::string& comment, const std::string& seq, std::string& name_state)
{
    std::stringstream ss;
    ss << comment;
    ss >> name_state;                                       /* a blank comment keeps the previous token */
    std::string sp = name_state.substr(1);
    std::string s(seq);
    for (auto& c : s) c = norm_base(c);
    db.id[sp] = db.name.size();
    db.name.push_back(sp);
    db.seq.push_back(s);
}

/* seq/SeqHelper.cpp:8-99: type from the first character, 4-line FASTQ, multi-line FASTA */
void load_seqs(SeqDB& db, const std::string& path)
{
    std::ifstream in(path);
    std::string first, name_state;
    bool fasta = false;
    {
        std::ifstream t(path);
        if (t && std::getline(t, first) && !first.empty()) fasta = first[0] == '>' || first[0] == ';';
    }
    if (!in.is_open()) return;
    if (fasta) {
        std::string line, name, buf;
        while (std::getline(in, line)) {
            if (!line.empty() && line[0] == '>') {
                if (!name.empty()) { add_seq(db, name, buf, name_state); buf.clear(); }
                name = line;
            } else {
                buf += line;
            }
        }
        if (!name.empty()) add_seq(db, name, buf, name_state);
    } else {
        std::string l[4];
        for (;;) {
            int i = 0;
            for (; i < 4; ++i) if (!std::getline(in, l[i])) break;
            if (i < 4) break;
            add_seq(db, l[0], l[1], name_state);
        }
    }
}

struct Aln {                                                /* align/AlignInf.hpp:18-45 */
    std::string q, r;
    size_t score, qb, qe, rb, re;
    bool fwd;
    std::vector<bool> qd, rd;
};

/* align/ParseAlignTools.cpp:8-26 */
void parse_diff(const std::string& a, const std::string& b, std::vector<bool>& qd, std::vector<bool>& rd)
{
    for (size_t i = 0; i < a.size(); ++i) {
        if (a[i] == '-') { qd.push_back(true); rd.push_back(false); }
        else if (b[i] == '-') { qd.push_back(false); rd.push_back(true); }
        else if (a[i] != b[i]) { qd.push_back(true); rd.push_back(true); }
        else { qd.push_back(false); rd.push_back(false); }
    }
}

bool by_score(const Aln& a, const Aln& b) { return a.score > b.score; }   /* align/AlignInf.cpp:31-33 */

/* align/MecatAlignDatabase.cpp:8-20 over align/AlignmentHelper.cpp:11-46: a header that does not parse still yields a
 * record (empty names, zeros) */
void load_mecat(std::vector<Aln>& out, const std::string& path)
{
    std::ifstream in(path);
    if (!in.is_open()) return;
    std::string l1, l2, l3;
    std::stringstream ss;
    std::string qn, rn, fw, sc;
    size_t qb = 0, qe = 0, qs = 0, rb = 0, re = 0, rs = 0;
    for (;;) {
        if (!std::getline(in, l1)) break;
        ss.clear();
        ss.str(l1);
        ss >> qn >> rn >> fw >> sc >> qb >> qe >> qs >> rb >> re >> rs;
        Aln a;
        if (!ss.fail()) {
            a.q = qn; a.r = rn; a.fwd = fw == "F"; a.score = (size_t)std::atoll(sc.c_str());
            a.qb = qb; a.qe = qe; a.rb = rb; a.re = re;
        } else {
            a.fwd = false; a.score = 0; a.qb = a.qe = a.rb = a.re = 0;
        }
        if (!std::getline(in, l2)) break;
        if (!std::getline(in, l3)) break;
        parse_diff(l2, l3, a.qd, a.rd);
        out.push_back(a);
    }
    std::sort(out.begin(), out.end(), by_score);
}

/* align/MummerAlignDatabaseV2.cpp:7-49: score = query span, bad headers drop the record */
void load_mummer(std::vector<Aln>& out, const std::string& path)
{
    std::ifstream in(path);
    if (in.is_open()) {
        std::string line, qn, rn, fw, ign, qline;
        std::stringstream ss;
        size_t qb = 0, qe = 0, rb = 0, re = 0;
        bool bad = false;
        for (size_t n = 0; std::getline(in, line); ++n) {
            if (n % 3 == 0) {
                ss.clear();
                ss.str(line);
                ss >> qn >> rn >> fw >> ign >> qb >> qe >> ign >> rb >> re;
                if (ss.fail()) bad = true;
            } else if (n % 3 == 1) {
                if (!bad) qline = line;
            } else {
                if (!bad) {
                    Aln a;
                    a.q = qn; a.r = rn; a.score = qe - qb; a.qb = qb; a.qe = qe; a.rb = rb; a.re = re; a.fwd = fw == "F";
                    parse_diff(qline, line, a.qd, a.rd);
                    out.push_back(a);
                }
                bad = false;
            }
        }
    }
    std::sort(out.begin(), out.end(), by_score);
}

struct ExAln { const Aln* a; size_t ref; };
bool ex_by_score(const ExAln& x, const ExAln& y) { return x.a->score > y.a->score; }

/* Aligner::mergeAlignInfHelper, align/Aligner.cpp:32-56 */
void group(std::vector<std::vector<ExAln>>& g, const std::vector<Aln>& alns, const SeqDB& qdb, const SeqDB& rdb)
{
    g.assign(qdb.name.size(), {});
    for (auto& a : alns)
        if (qdb.has(a.q) && rdb.has(a.r)) g[qdb.id.at(a.q)].push_back({&a, rdb.id.at(a.r)});
    for (auto& v : g) std::sort(v.begin(), v.end(), ex_by_score);
}

/* ParseAlignTools::exactAlign, align/ParseAlignTools.tcc:46-70 */
template <typename F>
void exact_align(size_t qb, size_t rb, bool forward, const std::vector<bool>& qd, const std::vector<bool>& rd, F f)
{
    if (qd.empty()) return;
    size_t cr = rb, cq = qb;
    for (size_t j = 0; j < qd.size(); ++j) {
        size_t idx = forward ? j : qd.size() - j - 1;
        if (!(qd[idx] ^ rd[idx])) { f(cq, cr); ++cr; ++cq; }
        else if (qd[idx]) ++cr;
        else { f(cq, cr); ++cq; }
    }
}

void flip(size_t& l, size_t& r, size_t len) { size_t t = l; l = len - r; r = len - t; }   /* Aligner.cpp:235-239 */

/* position/PositionMapper.cpp:8-47 */
struct Mapper {
    std::vector<size_t> start, size;
    explicit Mapper(const SeqDB& db)
    {
        for (auto& s : db.seq) size.push_back(s.size());
        if (size.empty()) return;
        start.push_back(size[0]);
        for (size_t i = 1; i < size.size(); ++i) start.push_back(start.back() + 3 * size[i - 1] + std::max(size[i - 1], size[i]));
        start.push_back(start.back() + 4 * size.back());
    }
    size_t single(int64_t idx, int64_t pos) const
    {
        if (idx == 0) return 0;
        int64_t i = idx > 0 ? idx - 1 : -idx - 1;
        size_t off = idx > 0 ? 0 : 2 * size[i];
        return start[i] + off + pos;
    }
};

struct Vertex {                                             /* node/KMerAdjNode.hpp:16-23 */
    std::vector<std::pair<uint64_t, int>> child;
    std::vector<DualPos> pos;
    std::vector<uint16_t> cnt;
};

/* KMerAdjNode::cluster, node/KMerAdjNode.tcc:74-111 with the predicate of PABruijnGraph::mergeKmerPosition,
 * graph/PABruijnGraph.cpp:259-274, and isPosSimilar :379-383.  Only `pos` shrinks; `cnt` keeps its length. */
bool similar(const DualPos& l, const DualPos& r, size_t dev)
{
    bool s1 = l.first != 0 && r.first != 0 && (size_t)(std::max(l.first, r.first) - std::min(l.first, r.first)) <= dev;
    bool s2 = l.second != 0 && r.second != 0 && (size_t)(std::max(l.second, r.second) - std::min(l.second, r.second)) <= dev;
    s1 = s1 || (l.first == 0 && r.first == 0);
    s2 = s2 || (l.second == 0 && r.second == 0);
    return s1 && s2;
}

void cluster(Vertex& v, size_t dev)
{
    size_t p = 0;
    for (size_t i = 0; i < v.pos.size() && i < v.cnt.size(); ++i) {
        DualPos item = v.pos[i];
        uint16_t c = v.cnt[i];
        bool sim = false;
        for (size_t j = 0; j < p; ++j)
            if (similar(item, v.pos[j], dev)) { sim = true; v.cnt[j] = (uint16_t)(v.cnt[j] + c); break; }
        if (!sim) { v.pos[p] = item; v.cnt[p] = c; ++p; }
    }
    v.pos.resize(p);
}

struct Graph {
    size_t k;
    std::vector<uint64_t> codes;                            /* _kmerIndexArr */
    std::unordered_map<uint64_t, uint64_t> index;           /* _kmerIndexMap */
    std::vector<Vertex> v;
};

/* kmer/FileKmerIterator.cpp:11-44 (iterate() re-reads the file from byte 0, so the header word k is one of the
 * "k-mers") and PABruijnGraph::PABruijnGraph, graph/PABruijnGraph.cpp:10-45 */
void load_graph(Graph& g, const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    g.k = 0;
    in.read(reinterpret_cast<char*>(&g.k), sizeof(size_t));
    std::ifstream in2(path, std::ios::binary);
    uint64_t w;
    while (in2.read(reinterpret_cast<char*>(&w), sizeof w)) g.codes.push_back(w);
    std::sort(g.codes.begin(), g.codes.end());
    g.codes.erase(std::unique(g.codes.begin(), g.codes.end()), g.codes.end());
    for (size_t i = 0; i < g.codes.size(); ++i) g.index[g.codes[i]] = i;
    g.v.resize(g.codes.size());
}

/* kmer/KmerHelper.cpp:7-25 */
void kmer_codes(std::vector<uint64_t>& codes, const std::string& s, size_t k)
{
    uint64_t code = 0, mask = (1UL << (k * 2)) - 1;
    auto acgt = [](char c) -> uint64_t { return c == 'C' || c == 'c' ? 1 : c == 'G' || c == 'g' ? 2 : c == 'T' || c == 't' ? 3 : 0; };
    for (size_t i = 0; i < k && i < s.size(); ++i) code = (code << 2) | acgt(s[i]);
    if (s.size() >= k) codes.push_back(code);
    for (size_t i = k; i < s.size(); ++i) { code = ((code << 2) | acgt(s[i])) & mask; codes.push_back(code); }
}

/* PABruijnGraph::addPositionAndEdge, graph/PABruijnGraph.cpp:238-257; sampleSequence, graph/PABruijnGraph.tcc:6-27 */
void add_position_and_edge(Graph& g, const std::string& seq, const std::vector<std::vector<DualPos>>& lists, size_t outer)
{
    std::vector<std::pair<uint64_t, size_t>> samples;
    std::vector<uint64_t> codes;
    kmer_codes(codes, seq, g.k);
    int64_t last = -1;
    if (seq.size() >= g.k) {
        for (size_t i = 0; i < seq.size() - g.k + 1; ++i) {
            if (lists[i].empty()) continue;
            auto it = g.index.find(codes[i]);
            if (it == g.index.end()) continue;
            if (last < 0 || i - (size_t)last >= outer) { samples.emplace_back(it->second, i); last = (int64_t)i; }
        }
    }
    for (auto& s : samples) {
        Vertex& v = g.v[s.first];
        v.pos.insert(v.pos.end(), lists[s.second].begin(), lists[s.second].end());
        v.cnt.insert(v.cnt.end(), lists[s.second].size(), (uint16_t)1);
    }
    for (size_t i = 1; i < samples.size(); ++i)
        g.v[samples[i - 1].first].child.emplace_back(samples[i].first, (int)(samples[i].second - samples[i - 1].second));
}

std::string revcomp(const std::string& s)                   /* seq/CompressedSeq.cpp:57-75, table "TGCA" */
{
    std::string r(s.size(), 'A');
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[i];
        r[s.size() - 1 - i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
    }
    return r;
}

struct Cfg {
    std::string ref, reads, ctgAln, refAln;
    std::vector<std::pair<std::string, bool>> contigs;
};

std::vector<Cfg> load_cfg(const std::string& path)          /* PGM/pagraph.cpp:29-49 */
{
    std::vector<Cfg> out;
    std::ifstream in(path);
    std::string line;
    while (std::getline(in, line)) {
        Cfg c;
        c.ref = line;
        std::getline(in, c.reads);
        std::getline(in, c.ctgAln);
        std::getline(in, c.refAln);
        while (std::getline(in, line) && !line.empty()) {
            c.contigs.emplace_back(line, false);
            std::getline(in, line);
            std::stringstream(line) >> c.contigs.back().second;
        }
        out.push_back(c);
    }
    return out;
}

typedef std::vector<std::vector<std::pair<RefPos, RefPos>>> PosLists;

void one_config(Graph& g, const Cfg& cfg, const SeqDB& ctgs, const SeqDB& refs, const std::vector<Aln>& c2r,
                const std::string& dir, size_t eps, size_t cov)
{
    for (auto& v : g.v) v = Vertex();                       /* resetAllNodes, PGM/pagraph.cpp:168 */
    SeqDB reads;
    load_seqs(reads, dir + "/" + cfg.reads);
    std::vector<Aln> r2c, r2r;
    load_mecat(r2c, dir + "/" + cfg.ctgAln);
    load_mecat(r2r, dir + "/" + cfg.refAln);

    const double readToCtgRatio = 0.35, readToRefRatio = 0.10;    /* PGM/pagraph.cpp:110-125 */
    const size_t outer = 3;
    Mapper cm(ctgs), rm(refs);

    /* Aligner::Aligner -> mergeAlignInf, align/Aligner.cpp:89-94 */
    std::vector<std::vector<ExAln>> readToCtg, readToRef, ctgToRef;
    group(readToCtg, r2c, reads, ctgs);
    group(readToRef, r2r, reads, refs);
    group(ctgToRef, c2r, ctgs, refs);
    std::vector<std::vector<size_t>> refCov(refs.seq.size());     /* covInfHelper :58-87 */
    for (size_t i = 0; i < refs.seq.size(); ++i) refCov[i].assign(refs.seq[i].size(), 0);
    for (auto& a : r2r) {
        if (!refs.has(a.r)) continue;
        auto& c = refCov[refs.id.at(a.r)];
        for (size_t j = a.rb; j < a.re; ++j) { if (j >= c.size()) break; ++c[j]; }
    }
    for (auto& c : refCov) std::sort(c.begin(), c.end());

    /* filters, PGM/pagraph.cpp:205-217 */
    std::vector<bool> refFlag(refs.seq.size(), false), ctgFlag(ctgs.seq.size(), false), ctgFwd(ctgs.seq.size(), true);
    if (refs.has(cfg.ref)) refFlag[refs.id.at(cfg.ref)] = true;
    for (auto& c : cfg.contigs)
        if (ctgs.has(c.first)) { ctgFlag[ctgs.id.at(c.first)] = true; ctgFwd[ctgs.id.at(c.first)] = c.second; }

    /* Aligner::simpleAlign, align/Aligner.cpp:96-201; AlignReference::insert, align/AlignReference.cpp:42-58 */
    std::vector<std::vector<std::vector<RefPos>>> fPos(ctgs.seq.size()), rPos(ctgs.seq.size());
    for (size_t i = 0; i < ctgs.seq.size(); ++i) { fPos[i].resize(ctgs.seq[i].size()); rPos[i].resize(ctgs.seq[i].size()); }
    for (size_t ci = 0; ci < ctgs.seq.size(); ++ci) {
        if (!ctgFlag[ci]) continue;
        for (auto& ex : ctgToRef[ci]) {
            if (!refFlag[ex.ref]) continue;
            bool forward = ex.a->fwd;
            if (ctgFwd[ci] != forward) continue;
            size_t cb = ex.a->qb, ce = ex.a->qe, len = ctgs.seq[ci].size();
            if (!forward) flip(cb, ce, len);
            std::vector<int64_t> refPos;
            exact_align(cb, ex.a->rb, true, ex.a->qd, ex.a->rd, [&](size_t, size_t r) { refPos.push_back((int64_t)r); });
            auto& tab = forward ? fPos[ci] : rPos[ci];
            for (size_t i = cb; i < ce; ++i) {
                if (i >= tab.size() || i - cb >= refPos.size()) break;   /* the reference reads out of bounds here */
                tab[i].emplace_back((int64_t)(ex.ref + 1), refPos[i - cb]);
            }
        }
    }
    /* Aligner::addExtraPosition :203-209, AlignReference::addExtraPosition :70-80 */
    for (size_t ci = 0; ci < ctgs.seq.size(); ++ci)
        if (ctgFlag[ci])
            for (auto& p : (ctgFwd[ci] ? fPos[ci] : rPos[ci]))
                if (p.empty()) p.emplace_back((int64_t)0, (int64_t)0);

    /* the functor of PositionProcessor::process, position/PositionProcessor.cpp:84-110; transformPosition :37-55 */
    auto feed = [&](size_t ri, const PosLists& fp, const PosLists& bp, bool fu, bool bu) {
        for (int s = 0; s < 2; ++s) {
            if (!(s == 0 ? fu : bu)) continue;
            const PosLists& p1 = s == 0 ? fp : bp;
            std::vector<std::vector<DualPos>> p2(p1.size());
            for (size_t i = 0; i < p1.size(); ++i)
                for (auto& pos : p1[i])
                    p2[i].emplace_back((uint32_t)cm.single(pos.first.first, pos.first.second),
                                       (uint32_t)rm.single(pos.second.first, pos.second.second));
            add_position_and_edge(g, s == 0 ? reads.seq[ri] : revcomp(reads.seq[ri]), p2, outer);
        }
    };

    /* Aligner::parseToCtg, align/Aligner.tcc:24-103 */
    for (size_t ri = 0; ri < reads.seq.size(); ++ri) {
        size_t readLen = reads.seq[ri].size();
        bool useful[2] = {false, false};
        PosLists fwdP(readLen), revP(readLen);
        for (auto& ex : readToCtg[ri]) {
            size_t ci = ex.ref;
            if (!ctgFlag[ci]) continue;
            size_t rb = ex.a->qb, re = ex.a->qe;
            if ((re - rb) * 1.0 / readLen < readToCtgRatio) continue;
            bool isFwd = ex.a->fwd;
            size_t cb = ex.a->rb, ce = ex.a->re, clen = ctgs.seq[ci].size();
            if (ce >= clen || cb >= clen) continue;
            if (!isFwd) flip(rb, re, readLen);
            for (int ii = 0; ii < 2; ++ii) {
                if ((ii == 0) == ctgFwd[ci]) {
                    PosLists& positions = isFwd ? fwdP : revP;
                    exact_align(rb, cb, ii == 0, ex.a->qd, ex.a->rd, [&](size_t cr, size_t cc) {
                        if (cr >= rb && cr < positions.size()) {
                            /* Aligner::queryContig, align/Aligner.cpp:222-233; AlignReference::query :60-68 */
                            auto& tab = ii == 0 ? fPos[ci] : rPos[ci];
                            bool any = false;
                            if ((int64_t)cc >= 0 && cc < tab.size())
                                for (auto& rp : tab[cc]) {
                                    positions[cr].emplace_back(RefPos(ii == 0 ? (int64_t)ci + 1 : -(int64_t)ci - 1, (int64_t)cc), rp);
                                    any = true;
                                }
                            useful[isFwd ? 0 : 1] = useful[isFwd ? 0 : 1] || any;
                        }
                    });
                }
                isFwd = !isFwd;
                flip(rb, re, readLen);
                flip(cb, ce, clen);
            }
        }
        feed(ri, fwdP, revP, useful[0], useful[1]);
    }
    /* position/PositionProcessor.cpp:121-123 */
    auto merge_edges = [&]() {                              /* KMerAdjNode::removeDuplicate, node/KMerAdjNode.tcc:46-70 */
        for (auto& v : g.v) {
            if (v.child.empty()) continue;
            std::sort(v.child.begin(), v.child.end());
            v.child.erase(std::unique(v.child.begin(), v.child.end()), v.child.end());
        }
    };
    merge_edges();
    for (auto& v : g.v) cluster(v, eps);

    /* Aligner::parseToRef, align/Aligner.tcc:106-171 */
    for (size_t ri = 0; ri < reads.seq.size(); ++ri) {
        size_t readLen = reads.seq[ri].size();
        bool useful[2] = {false, false};
        PosLists fwdP(readLen), revP(readLen);
        for (auto& ex : readToRef[ri]) {
            size_t fi = ex.ref;
            if (!refFlag[fi]) continue;
            size_t rb = ex.a->qb, re = ex.a->qe;
            if ((re - rb) * 1.0 / readLen < readToRefRatio) continue;
            bool isFwd = ex.a->fwd;
            size_t maxCov = 0;
            for (size_t p = ex.a->rb; p < ex.a->re; ++p) {
                if (p >= refCov[fi].size()) break;
                maxCov = std::max(maxCov, refCov[fi][p]);
            }
            if (maxCov < cov) continue;
            if (!isFwd) flip(rb, re, readLen);
            PosLists& positions = isFwd ? fwdP : revP;
            useful[isFwd ? 0 : 1] = true;
            exact_align(rb, ex.a->rb, true, ex.a->qd, ex.a->rd, [&](size_t cr, size_t cf) {
                if (cr >= rb && cr < positions.size())
                    positions[cr].emplace_back(RefPos(0, 0), RefPos((int64_t)fi + 1, (int64_t)cf));
            });
        }
        feed(ri, fwdP, revP, useful[0], useful[1]);
    }
    /* position/PositionProcessor.cpp:135-139 */
    merge_edges();
    for (auto& v : g.v) cluster(v, eps);
    for (auto& v : g.v) {                                   /* sortWithCount, node/KMerAdjNode.tcc:115-136 */
        std::vector<std::pair<DualPos, uint16_t>> t;
        for (size_t i = 0; i < v.pos.size() && i < v.cnt.size(); ++i) t.emplace_back(v.pos[i], v.cnt[i]);
        std::sort(t.begin(), t.end(),
                  [](const std::pair<DualPos, uint16_t>& a, const std::pair<DualPos, uint16_t>& b) { return a.first < b.first; });
        v.pos.clear();
        v.cnt.clear();
        for (auto& p : t) { v.pos.push_back(p.first); v.cnt.push_back(p.second); }
    }
}

}  // namespace

extern "C" int ag2o_pagraph_dump(const char* kmer, const char* ctg, const char* ref, const char* dir, const char* aln,
                                 long eps, long cov, const char* out_path)
{
    FILE* out = fopen(out_path, "w");
    if (!out) return -1;
    Graph g;
    load_graph(g, kmer);
    SeqDB ctgs, refs;
    load_seqs(ctgs, ctg);
    load_seqs(refs, ref);
    std::vector<Aln> c2r;
    load_mummer(c2r, aln);
    auto cfgs = load_cfg(std::string(dir) + "/config.txt");
    int n = 0;
    for (auto& cfg : cfgs) {
        one_config(g, cfg, ctgs, refs, c2r, dir, (size_t)eps, (size_t)cov);
        fprintf(out, "#config %d %s\n", n++, cfg.ref.c_str());
        for (size_t v = 0; v < g.v.size(); ++v) {
            auto& x = g.v[v];
            if (x.pos.empty() && x.child.empty()) continue;
            fprintf(out, "V %zu %llu P %zu", v, (unsigned long long)g.codes[v], x.pos.size());
            for (size_t i = 0; i < x.pos.size(); ++i) fprintf(out, " %u,%u,%u", x.pos[i].first, x.pos[i].second, (unsigned)x.cnt[i]);
            fprintf(out, " E %zu", x.child.size());
            for (auto& e : x.child) fprintf(out, " %llu,%d", (unsigned long long)e.first, e.second);
            fputc('\n', out);
        }
    }
    fclose(out);
    return 0;
}
