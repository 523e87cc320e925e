// pg_job.cpp -- the file handling of `pagraph` (PAGraph/src/main/pagraph.cpp:127-204) over the C ABI of
// include/ag2_pagraph.h: solid k-mer file, contig / reference / read databases, the three `.ref` alignment files and
// <pre dir>/config.txt.  Host C++ like the reference; it parses headers and names, the alignment lines go to the GPU as
// the file bytes they are.  Used by the drop-in executable (host/pagraph_main.cpp), tests and bench.py.
// Reference paths are relative to PAGraph/src/tools/.
#include "../../include/ag2_b200.h"
#include "../../include/ag2_pagraph.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

bool read_file(const std::string& path, std::string& out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? fread(&out[0], 1, (size_t)n, f) : 0;
    fclose(f);
    out.resize(got);
    return true;
}

// std::getline over a memory buffer: [b, e) of the next line, false at the end
struct Lines {
    const std::string& s;
    size_t at = 0;
    explicit Lines(const std::string& str) : s(str) {}
    bool next(size_t& b, size_t& e)
    {
        if (at >= s.size()) return false;
        b = at;
        const void* nl = memchr(s.data() + at, '\n', s.size() - at);
        e = nl ? (size_t)((const char*)nl - s.data()) : s.size();
        at = e + 1;
        return true;
    }
};

struct SeqDB {                                   // seq/AutoSeqDatabase.cpp:9-22, seq/SeqHelper.cpp:8-99
    std::vector<std::string> name;
    std::vector<int64_t> off{0};
    std::string bases;
    std::unordered_map<std::string, int32_t> id;
    int64_t size() const { return (int64_t)name.size(); }
    int64_t len(int64_t i) const { return off[i + 1] - off[i]; }
    int32_t find(const std::string& n) const
    {
        auto it = id.find(n);
        return it == id.end() ? -1 : it->second;
    }
};

void first_token(const std::string& s, size_t b, size_t e, std::string& tok)   // `ss >> name`: keeps the old value if blank
{
    while (b < e && isspace((unsigned char)s[b])) ++b;
    size_t t = b;
    while (t < e && !isspace((unsigned char)s[t])) ++t;
    if (t > b) tok.assign(s, b, t - b);
}

bool load_seqs(const std::string& path, SeqDB& db)
{
    std::string buf;
    if (!read_file(path, buf)) return false;
    Lines ln(buf);
    size_t b, e;
    std::string tok;
    auto add = [&](size_t nb, size_t ne) {
        first_token(buf, nb, ne, tok);
        const std::string sp = tok.empty() ? tok : tok.substr(1);
        db.id[sp] = (int32_t)db.name.size();
        db.name.push_back(sp);
        db.off.push_back((int64_t)db.bases.size());
    };
    const bool fasta = !buf.empty() && (buf[0] == '>' || buf[0] == ';');   // SeqHelper::testFileType
    if (fasta) {                                  // SeqHelper::loadFromFasta: a record closes when the next header arrives;
        bool have = false;                        // lines before the first header join the first record
        size_t nb = 0, ne = 0;
        while (ln.next(b, e)) {
            if (e > b && buf[b] == '>') {
                if (have) add(nb, ne);
                nb = b;
                ne = e;
                have = true;
            } else {
                db.bases.append(buf, b, e - b);
            }
        }
        if (have) add(nb, ne);
    } else {
        size_t lb[4], le[4];
        for (;;) {
            int i = 0;
            for (; i < 4; ++i)
                if (!ln.next(lb[i], le[i])) break;
            if (i < 4) break;
            db.bases.append(buf, lb[1], le[1] - lb[1]);
            add(lb[0], le[0]);
        }
    }
    return true;
}

struct AlnFile {
    std::string text;                 // the file
    std::vector<ag2_pg_aln> rec;
};

size_t split(const std::string& s, size_t b, size_t e, size_t* tb, size_t* te, size_t max_tok)
{
    size_t n = 0;
    while (b < e && n < max_tok) {
        while (b < e && isspace((unsigned char)s[b])) ++b;
        if (b >= e) break;
        tb[n] = b;
        while (b < e && !isspace((unsigned char)s[b])) ++b;
        te[n++] = b;
    }
    return n;
}

bool parse_size(const std::string& s, size_t b, size_t e, uint64_t& v)   // `ss >> size_t` on a clean token
{
    if (b >= e) return false;
    bool neg = false;
    if (s[b] == '+' || s[b] == '-') { neg = s[b] == '-'; ++b; }
    if (b >= e) return false;
    uint64_t x = 0;
    for (; b < e; ++b) {
        if (s[b] < '0' || s[b] > '9') return false;
        x = x * 10 + (uint64_t)(s[b] - '0');
    }
    v = neg ? (uint64_t)0 - x : x;
    return true;
}

// align/AlignmentHelper.cpp:11-46 + align/MecatAlignDatabase.cpp:8-20 (mummer = false): every 3-line group is a record,
// a header that does not parse gives an anonymous record with score 0.
// align/MummerAlignDatabaseV2.cpp:7-49 (mummer = true): score = query span, bad headers drop the record.
bool load_alns(const std::string& path, bool mummer, const SeqDB& qdb, const SeqDB& tdb, AlnFile& f)
{
    if (!read_file(path, f.text)) return true;   // a missing file is an empty database in the reference too
    const std::string& s = f.text;
    Lines ln(s);
    size_t b[3], e[3], tb[10], te[10];
    std::string qn, tn;
    for (;;) {
        int i = 0;
        for (; i < 3; ++i)
            if (!ln.next(b[i], e[i])) break;
        if (i < 3) break;
        ag2_pg_aln a{};
        a.query = a.target = -1;
        bool ok;
        uint64_t v[6] = {0, 0, 0, 0, 0, 0};
        if (!mummer) {
            ok = split(s, b[0], e[0], tb, te, 10) == 10;
            const int num[6] = {4, 5, 6, 7, 8, 9};
            for (int j = 0; ok && j < 6; ++j) ok = parse_size(s, tb[num[j]], te[num[j]], v[j]);
            if (ok) {
                a.score = (uint64_t)std::atoll(std::string(s, tb[3], te[3] - tb[3]).c_str());
                a.qb = (int64_t)v[0]; a.qe = (int64_t)v[1]; a.tb = (int64_t)v[3]; a.te = (int64_t)v[4];
            }
        } else {
            ok = split(s, b[0], e[0], tb, te, 9) == 9;
            const int num[4] = {4, 5, 7, 8};
            for (int j = 0; ok && j < 4; ++j) ok = parse_size(s, tb[num[j]], te[num[j]], v[j]);
            if (!ok) continue;
            a.qb = (int64_t)v[0]; a.qe = (int64_t)v[1]; a.tb = (int64_t)v[2]; a.te = (int64_t)v[3];
            a.score = v[1] - v[0];
        }
        if (ok) {
            qn.assign(s, tb[0], te[0] - tb[0]);
            tn.assign(s, tb[1], te[1] - tb[1]);
            a.query = qdb.find(qn);
            a.target = tdb.find(tn);
            a.forward = te[2] - tb[2] == 1 && s[tb[2]] == 'F';
        }
        const size_t l2 = e[1] - b[1], l3 = e[2] - b[2];
        a.ncols = (int32_t)(l2 < l3 ? l2 : l3);   // the reference indexes line 3 by line 2's length (undefined when shorter)
        a.q_off = (int64_t)b[1];
        a.t_off = (int64_t)b[2];
        f.rec.push_back(a);
    }
    return true;
}

struct Block {                                   // one block of <pre dir>/config.txt, PGM/pagraph.cpp:21-49
    std::string ref, reads, ctg_aln, ref_aln;
    std::vector<std::pair<std::string, bool>> contigs;
};

std::vector<Block> load_cfg(const std::string& path)
{
    std::vector<Block> out;
    std::ifstream in(path);
    std::string line;
    while (std::getline(in, line)) {
        Block c;
        c.ref = line;
        std::getline(in, c.reads);
        std::getline(in, c.ctg_aln);
        std::getline(in, c.ref_aln);
        while (std::getline(in, line) && !line.empty()) {
            c.contigs.emplace_back(line, false);
            std::getline(in, line);
            std::stringstream(line) >> c.contigs.back().second;
        }
        out.push_back(c);
    }
    return out;
}

}  // namespace

struct ag2_pg_job {
    ag2_pg* pg = nullptr;
    std::string dir, err;
    SeqDB ctgs, refs, reads;
    AlnFile c2r, r2c, r2r;
    std::vector<Block> blocks;
    std::vector<uint64_t> codes;
    int64_t first_read = 0, n_local = 0;
    int32_t k = 0;                               // FileKmerIterator::kSize: the first word of the k-mer file
    std::unordered_set<std::string> ok_ctg;      // okCtg of PGM/pagraph.cpp:164,259-261 (same container: contig.txt is written in its order)
};

namespace {
int jfail(ag2_pg_job* j, int code, const std::string& what)
{
    j->err = what;
    if (j->pg && code != AG2_EINVAL && *ag2_pg_last_error(j->pg)) j->err += std::string(": ") + ag2_pg_last_error(j->pg);
    return code;
}
}  // namespace

extern "C" {

int ag2_pg_job_open(int device, const char* kmer_path, const char* ctg_path, const char* ref_path, const char* pre_dir,
                    const char* ctg_to_ref_path, ag2_pg_job** out)
{
    if (!out || !kmer_path || !ctg_path || !ref_path || !pre_dir || !ctg_to_ref_path) return AG2_EINVAL;
    *out = nullptr;
    ag2_pg_job* j = new (std::nothrow) ag2_pg_job();
    if (!j) return AG2_ENOMEM;
    *out = j;   // returned even on failure so that ag2_pg_job_error can explain; the caller closes it
    int rc = ag2_pg_create(device, &j->pg);
    if (rc != AG2_OK) return jfail(j, rc, "no usable CUDA device (aligngraph2_b200 has no CPU path)");
    j->dir = pre_dir;
    std::string kbuf;
    if (!read_file(kmer_path, kbuf) || kbuf.size() < 8) return jfail(j, AG2_EINVAL, std::string("cannot read ") + kmer_path);
    std::vector<uint64_t> words(kbuf.size() / 8);
    memcpy(words.data(), kbuf.data(), words.size() * 8);
    j->k = (int32_t)words[0];
    int64_t nv = 0;
    if ((rc = ag2_pg_set_kmers(j->pg, words.data(), (int64_t)words.size(), &nv)) != AG2_OK) return jfail(j, rc, "ag2_pg_set_kmers");
    j->codes.resize((size_t)nv);
    if ((rc = ag2_pg_fetch_codes(j->pg, j->codes.data(), nv)) != AG2_OK) return jfail(j, rc, "ag2_pg_fetch_codes");
    if (!load_seqs(ctg_path, j->ctgs)) return jfail(j, AG2_EINVAL, std::string("cannot read ") + ctg_path);
    if (!load_seqs(ref_path, j->refs)) return jfail(j, AG2_EINVAL, std::string("cannot read ") + ref_path);
    std::vector<int64_t> cl((size_t)j->ctgs.size()), rl((size_t)j->refs.size());
    for (int64_t i = 0; i < j->ctgs.size(); ++i) cl[i] = j->ctgs.len(i);
    for (int64_t i = 0; i < j->refs.size(); ++i) rl[i] = j->refs.len(i);
    if ((rc = ag2_pg_set_targets(j->pg, cl.data(), (int64_t)cl.size(), rl.data(), (int64_t)rl.size())) != AG2_OK) return jfail(j, rc, "ag2_pg_set_targets");
    load_alns(ctg_to_ref_path, true, j->ctgs, j->refs, j->c2r);
    j->blocks = load_cfg(j->dir + "/config.txt");
    return AG2_OK;
}

void ag2_pg_job_close(ag2_pg_job* j)
{
    if (!j) return;
    ag2_pg_destroy(j->pg);
    delete j;
}

const char* ag2_pg_job_error(const ag2_pg_job* j) { return j ? j->err.c_str() : "null job"; }
int ag2_pg_job_blocks(const ag2_pg_job* j) { return j ? (int)j->blocks.size() : 0; }
ag2_pg* ag2_pg_job_handle(ag2_pg_job* j) { return j ? j->pg : nullptr; }
const char* ag2_pg_job_block_ref(const ag2_pg_job* j, int block)
{
    return j && block >= 0 && block < (int)j->blocks.size() ? j->blocks[block].ref.c_str() : "";
}

// one config block: the body of the loop at PGM/pagraph.cpp:167-218 up to pp.preProcess().  rank / world split the
// read database into contiguous ranges (world = 1: everything).
int ag2_pg_job_load_block(ag2_pg_job* j, int block, int rank, int world)
{
    if (!j || !j->pg || block < 0 || block >= (int)j->blocks.size() || world < 1 || rank < 0 || rank >= world) return AG2_EINVAL;
    const Block& B = j->blocks[block];
    j->reads = SeqDB();
    j->r2c = AlnFile();
    j->r2r = AlnFile();
    if (!load_seqs(j->dir + "/" + B.reads, j->reads)) return jfail(j, AG2_EINVAL, "cannot read " + j->dir + "/" + B.reads);
    const int64_t n = j->reads.size();
    const int64_t first = n * rank / world, last = n * (rank + 1) / world;
    j->first_read = first;
    j->n_local = last - first;
    std::vector<int64_t> offs((size_t)(last - first) + 1);
    for (int64_t i = first; i <= last; ++i) offs[i - first] = j->reads.off[i] - j->reads.off[first];
    int rc = ag2_pg_set_reads(j->pg, j->reads.bases.data() + j->reads.off[first], offs.data(), last - first, first);
    if (rc != AG2_OK) return jfail(j, rc, "ag2_pg_set_reads");
    rc = ag2_pg_set_alignments(j->pg, AG2_PG_CTG_TO_REF, j->c2r.rec.data(), (int64_t)j->c2r.rec.size(), j->c2r.text.data(), (int64_t)j->c2r.text.size());
    if (rc != AG2_OK) return jfail(j, rc, "ag2_pg_set_alignments(ctg->ref)");
    load_alns(j->dir + "/" + B.ctg_aln, false, j->reads, j->ctgs, j->r2c);
    load_alns(j->dir + "/" + B.ref_aln, false, j->reads, j->refs, j->r2r);
    AlnFile* files[2] = {&j->r2c, &j->r2r};
    for (int w = 0; w < 2; ++w) {
        AlnFile& f = *files[w];
        // the text this rank needs: the byte range spanned by the records of its own reads
        int64_t lo = (int64_t)f.text.size(), hi = 0;
        for (auto& a : f.rec) {
            if (a.query < first || a.query >= last) { a.ncols = 0; a.q_off = a.t_off = 0; continue; }
            lo = std::min(lo, std::min(a.q_off, a.t_off));
            hi = std::max(hi, std::max(a.q_off, a.t_off) + a.ncols);
        }
        if (hi <= lo) lo = hi = 0;
        for (auto& a : f.rec)
            if (a.query >= first && a.query < last) { a.q_off -= lo; a.t_off -= lo; }
        rc = ag2_pg_set_alignments(j->pg, w, f.rec.data(), (int64_t)f.rec.size(), f.text.data() + lo, hi - lo);
        if (rc != AG2_OK) return jfail(j, rc, "ag2_pg_set_alignments");
    }
    // pp.clearRefFilter(false); pp.clearCtgFilter(false); pp.setRefFilter(config.ref, true); pp.setCtgFilter(...)
    std::vector<uint8_t> rf((size_t)j->refs.size() + 1, 0), cf((size_t)j->ctgs.size() + 1, 0), cw((size_t)j->ctgs.size() + 1, 1);
    if (j->refs.find(B.ref) >= 0) rf[j->refs.find(B.ref)] = 1;
    for (auto& c : B.contigs) {
        const int32_t id = j->ctgs.find(c.first);
        if (id >= 0) { cf[id] = 1; cw[id] = c.second ? 1 : 0; }
    }
    if ((rc = ag2_pg_set_filters(j->pg, rf.data(), cf.data(), cw.data())) != AG2_OK) return jfail(j, rc, "ag2_pg_set_filters");
    return AG2_OK;
}

// the graph of the handle as text (the parity tests compare this dump with the same dump of the reference classes):
//   #config <block> <ref name>  then per non-empty vertex  V <idx> <code> P <n> {ctg,ref,count}.. E <m> {to,step}..
int ag2_pg_job_dump(ag2_pg_job* j, int block, const char* path, int append)
{
    if (!j || !j->pg || !path) return AG2_EINVAL;
    ag2_pg_stats st;
    ag2_pg_get_stats(j->pg, &st);
    const int64_t nv = st.n_vertices;
    std::vector<int64_t> po((size_t)nv + 1), eo((size_t)nv + 1);
    std::vector<uint32_t> ctg((size_t)st.positions + 1), ref((size_t)st.positions + 1), to((size_t)st.edges + 1);
    std::vector<uint16_t> cnt((size_t)st.positions + 1);
    std::vector<int32_t> step((size_t)st.edges + 1);
    int rc = ag2_pg_graph_fetch(j->pg, po.data(), ctg.data(), ref.data(), cnt.data(), st.positions, eo.data(), to.data(), step.data(), st.edges);
    if (rc != AG2_OK) return jfail(j, rc, "ag2_pg_graph_fetch");
    FILE* out = fopen(path, append ? "a" : "w");
    if (!out) return jfail(j, AG2_EINVAL, std::string("cannot write ") + path);
    fprintf(out, "#config %d %s\n", block, block >= 0 && block < (int)j->blocks.size() ? j->blocks[block].ref.c_str() : "");
    for (int64_t v = 0; v < nv; ++v) {
        if (po[v] == po[v + 1] && eo[v] == eo[v + 1]) continue;
        fprintf(out, "V %lld %llu P %lld", (long long)v, (unsigned long long)j->codes[v], (long long)(po[v + 1] - po[v]));
        for (int64_t i = po[v]; i < po[v + 1]; ++i) fprintf(out, " %u,%u,%u", ctg[i], ref[i], (unsigned)cnt[i]);
        fprintf(out, " E %lld", (long long)(eo[v + 1] - eo[v]));
        for (int64_t i = eo[v]; i < eo[v + 1]; ++i) fprintf(out, " %u,%d", to[i], step[i]);
        fputc('\n', out);
    }
    fclose(out);
    return AG2_OK;
}

// B9 on the graph of the handle: PAssembly::testTravel5 as run2 calls it (PGM/pagraph.cpp:244-261)
int ag2_pg_job_travel(ag2_pg_job* j, int block, const ag2_pg_travel_params* prm, const char* out_dir)
{
    if (!j || !j->pg || !prm || !out_dir || block < 0 || block >= (int)j->blocks.size()) return AG2_EINVAL;
    ag2_pg_stats st;
    ag2_pg_get_stats(j->pg, &st);
    const int64_t nv = st.n_vertices;
    std::vector<int64_t> po((size_t)nv + 1), eo((size_t)nv + 1);
    std::vector<uint32_t> ctg((size_t)st.positions + 1), ref((size_t)st.positions + 1), to((size_t)st.edges + 1);
    std::vector<uint16_t> cnt((size_t)st.positions + 1);
    std::vector<int32_t> step((size_t)st.edges + 1);
    int rc = ag2_pg_graph_fetch(j->pg, po.data(), ctg.data(), ref.data(), cnt.data(), st.positions, eo.data(), to.data(), step.data(), st.edges);
    if (rc != AG2_OK) return jfail(j, rc, "ag2_pg_graph_fetch");
    ag2_pg_graph_view g{j->k, nv, j->codes.data(), po.data(), ctg.data(), ref.data(), cnt.data(), eo.data(), to.data(), step.data()};
    std::vector<const char*> cn, rn;
    for (auto& n : j->ctgs.name) cn.push_back(n.c_str());
    for (auto& n : j->refs.name) rn.push_back(n.c_str());
    ag2_pg_seqs cs{j->ctgs.size(), cn.data(), j->ctgs.bases.data(), j->ctgs.off.data()};
    ag2_pg_seqs rs{j->refs.size(), rn.data(), j->refs.bases.data(), j->refs.off.data()};
    const Block& B = j->blocks[block];
    std::vector<int32_t> use;
    std::vector<uint8_t> fwd;
    for (auto& c : B.contigs) {
        const int32_t id = j->ctgs.find(c.first);
        if (id < 0) return jfail(j, AG2_EINVAL, "config.txt names a contig that is not in the contig file: " + c.first);
        use.push_back(id);
        fwd.push_back(c.second ? 1 : 0);
    }
    std::vector<int32_t> ok(use.size() * 2 + 1);
    int64_t n_ok = 0;
    const std::string prefix = std::to_string(block) + "_";
    rc = ag2_pg_travel(&g, &cs, &rs, use.data(), fwd.data(), (int64_t)use.size(), prm, out_dir, prefix.c_str(), ok.data(), &n_ok);
    if (rc != AG2_OK) return jfail(j, rc, std::string("ag2_pg_travel: cannot write under ") + out_dir);
    for (int64_t i = 0; i < n_ok; ++i) j->ok_ctg.emplace(j->ctgs.name[ok[i]]);
    return AG2_OK;
}

int ag2_pg_job_write_contig_list(ag2_pg_job* j, const char* out_dir)
{
    if (!j || !out_dir) return AG2_EINVAL;
    FILE* f = fopen((std::string(out_dir) + "/contig.txt").c_str(), "w");
    if (!f) return jfail(j, AG2_EINVAL, std::string("cannot write ") + out_dir + "/contig.txt");
    for (auto& n : j->ok_ctg) fprintf(f, "%s\n", n.c_str());
    fclose(f);
    return AG2_OK;
}

}  // extern "C"
