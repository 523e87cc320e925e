// kmer_counter_main.cpp -- drop-in replacement of PAGraph's kmer_counter executable:
//   kmer_counter -t N -i reads -o out.bin -k K [-m 0.2 -p 4 -s 10240]
// (PAGraph/src/main/kmer_counter.cpp:98-140; AlignGraph2.py:215-221).  The abundance table, the cut and the
// selection run on the GPU through the C ABI (include/ag2_b200.h).  Output: native `size_t k`, then the solid k-mer
// codes as uint64 in ascending order -- what the reference writes with -t 1 (with -t N it writes the same set in a
// thread-strided order; the only consumer, FileKmerIterator -> PABruijnGraph, sorts it).  -t, -p and -s are accepted
// and have no effect here.  No CPU fallback: without a GPU the program exits 1.
#include "../../include/ag2_b200.h"
#include "shard_split.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

namespace {

void usage() { std::cerr << "  kmer_counter {OPTIONS}\n\n  OPTIONS:\n      -h, --help  -t[thread_num]  -i[path]  -o[path]  -k[k]  -m[threshold]  -p[size]  -s[size]\n"; }

struct Args {
    unsigned threads = 16;
    std::string in, out;
    size_t k = 14, part = 4, size = 10240;
    double threshold = 0.2;
};

bool take(int argc, char **argv, int &i, const char *s, const char *l, std::string &val)
{
    const std::string a = argv[i];
    const std::string ls = std::string("--") + l, ss = std::string("-") + s;
    if (a == ss || a == ls) {
        if (i + 1 >= argc) throw std::runtime_error("Flag '" + a + "' requires an argument");
        val = argv[++i];
        return true;
    }
    if (a.rfind(ls + "=", 0) == 0) { val = a.substr(ls.size() + 1); return true; }
    if (a.rfind(ss, 0) == 0 && a.size() > 2 && a[1] != '-') { val = a.substr(2); return true; }
    return false;
}

void die(ag2_ctx *ctx, const char *what, int rc)
{
    fprintf(stderr, "kmer_counter (aligngraph2_b200): %s failed (%d): %s\n", what, rc, ag2_last_error(ctx));
    exit(1);
}

struct Batch {
    std::string bases;
    std::vector<int64_t> offs{0};
    void add(const std::string &s) { bases += s; offs.push_back((int64_t)bases.size()); }
    bool empty() const { return offs.size() == 1; }
    void clear() { bases.clear(); offs.assign(1, 0); }
};

} // namespace

int main(int argc, char **argv)
{
    Args a;
    if (argc <= 1) { usage(); return 0; }
    try {
        for (int i = 1; i < argc; ++i) {
            std::string v;
            const std::string s = argv[i];
            if (s == "-h" || s == "--help") { usage(); return 0; }
            else if (take(argc, argv, i, "t", "thread", v)) a.threads = (unsigned)std::stoul(v);
            else if (take(argc, argv, i, "i", "in", v)) a.in = v;
            else if (take(argc, argv, i, "o", "out", v)) a.out = v;
            else if (take(argc, argv, i, "k", "kmer", v)) a.k = std::stoul(v);
            else if (take(argc, argv, i, "m", "min", v)) a.threshold = std::stod(v);
            else if (take(argc, argv, i, "p", "part", v)) a.part = std::stoul(v);
            else if (take(argc, argv, i, "s", "size", v)) a.size = std::stoul(v);
            else throw std::runtime_error("Flag could not be matched: " + s);
        }
    } catch (const std::exception &e) {
        std::cerr << e.what() << std::endl;
        usage();
        return 1;
    }
    // The GPUs the counting runs on: all visible devices or the list AG2_DEVICES names; every read batch
    // goes to the next device in turn, each device counts into its own 4^k table, the tables are summed on the first
    // (ag2_kmer_merge, NVLink peer reads) before the cut.
    std::vector<int> devs;
    if (const char *e = getenv("AG2_DEVICES")) {
        devs = ag2host::parse_device_list(e);
    } else {
        int nd = 0;
        if (ag2_device_count(&nd) == AG2_OK)
            for (int d = 0; d < nd; ++d) devs.push_back(d);
    }
    std::vector<ag2_ctx *> ctxs;
    int rc = AG2_OK;
    for (int d : devs) {
        ag2_ctx *c = nullptr;
        if ((rc = ag2_ctx_create(d, &c)) == AG2_OK) ctxs.push_back(c);
    }
    if (ctxs.empty()) {
        fprintf(stderr, "kmer_counter (aligngraph2_b200): no usable CUDA device (%d); there is no CPU path\n", rc);
        return 1;
    }
    ag2_ctx *ctx = ctxs[0];
    for (ag2_ctx *c : ctxs)
        if ((rc = ag2_kmer_begin(c, (int)a.k)) != AG2_OK) die(c, "ag2_kmer_begin", rc);
    size_t turn = 0;

    // SeqHelper::autoLoadFromFile (PAGraph/src/tools/seq/SeqHelper.cpp:8-97): type from the first character
    Batch b;
    size_t batch_bytes = (size_t)1 << 30;
    if (const char *e = getenv("AG2_KMER_BATCH_BYTES")) batch_bytes = (size_t)std::max(1ll, atoll(e));   // test knob: many small batches
    auto flush = [&]() {
        if (b.empty()) return;
        ag2_ctx *c = ctxs[turn++ % ctxs.size()];
        if ((rc = ag2_reads_load(c, b.bases.data(), b.offs.data(), (int64_t)b.offs.size() - 1)) != AG2_OK) die(c, "ag2_reads_load", rc);
        if ((rc = ag2_kmer_add_reads(c)) != AG2_OK) die(c, "ag2_kmer_add_reads", rc);
        b.clear();
    };
    auto add = [&](const std::string &seq) {
        if (seq.empty()) return;
        b.add(seq);
        if (b.bases.size() >= batch_bytes) flush();
    };
    {
        std::ifstream in(a.in);
        std::string first;
        bool fasta = false;
        if (in && std::getline(in, first) && !first.empty()) fasta = first.front() == '>' || first.front() == ';';
        in.clear();
        in.seekg(0);
        if (in.is_open()) {
            std::string line;
            if (fasta) {
                std::string name, buffer;
                while (std::getline(in, line)) {
                    if (!line.empty() && line[0] == '>') {
                        if (!name.empty()) { add(buffer); buffer.clear(); }
                        name = line;
                    } else {
                        buffer += line;
                    }
                }
                if (!name.empty()) add(buffer);
            } else {
                std::string l2;
                for (size_t n = 0; std::getline(in, line); ++n) {
                    if (n % 4 == 1) l2 = line;
                    if (n % 4 == 3) add(l2);
                }
            }
        }
    }
    flush();
    for (size_t i = 1; i < ctxs.size(); ++i)
        if ((rc = ag2_kmer_merge(ctx, ctxs[i])) != AG2_OK) die(ctx, "ag2_kmer_merge", rc);
    int64_t cut = 0, n = 0;
    if ((rc = ag2_kmer_solid(ctx, a.threshold, &cut, &n)) != AG2_OK) die(ctx, "ag2_kmer_solid", rc);
    std::vector<uint64_t> codes((size_t)n + 1);
    if ((rc = ag2_kmer_fetch(ctx, codes.data(), n)) != AG2_OK) die(ctx, "ag2_kmer_fetch", rc);
    std::ofstream of(a.out, std::ios::binary);
    const size_t k = a.k;
    of.write(reinterpret_cast<const char *>(&k), sizeof(size_t));
    of.write(reinterpret_cast<const char *>(codes.data()), (std::streamsize)((size_t)n * sizeof(uint64_t)));
    of.close();
    for (ag2_ctx *c : ctxs) ag2_ctx_destroy(c);
    return 0;
}
