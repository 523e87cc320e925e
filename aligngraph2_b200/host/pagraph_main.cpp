// pagraph_main.cpp -- drop-in replacement of PAGraph's `pagraph` executable:
//   pagraph -t N -r dummy -k solid.bin -c ctg.fasta -R ref.fasta -p <pre dir> -a aln -o <out> [-l minLen] [--epsilon E] [-v V]
// (PAGraph/src/main/pagraph.cpp:69-272; AlignGraph2.py:414-427).  Per block of <pre dir>/config.txt: the A-Bruijn graph
// build (B2-B8) runs on the GPU through the C ABI of include/ag2_pagraph.h, the traversal (B9) on the host over the
// fetched graph; writes <out>/<block>_<ctg>_<0|1>.txt, .fasta / .help / .con and <out>/contig.txt like the reference.
// The graph is the one `pagraph -t 1` builds (the only deterministic setting), whatever -t says; -t keeps
// its other meaning: min(t, 8) start vertices per traversal round (PAlgorithm.cpp:146).
// A repeated flag keeps its last value, so the pipeline's second `-r <minLen>` lands in the unused read path as it does
// in the reference.  No CPU fallback: without a GPU the program exits 1.
#include "../../include/ag2_b200.h"
#include "../../include/ag2_pagraph.h"

#include "shard_split.h"

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

void usage()
{
    std::cerr << "  pagraph {OPTIONS}\n\n  OPTIONS:\n      -h, --help  -t[thread_num]  -k[path]  -r[path]  -c[path]  -R[path]  -p[path]  -a[path]  -o[path]"
                 "  -l[len]  --epsilon=[dist]  -v[cov]\n";
}

bool take(int argc, char** argv, int& i, const char* s, const char* l, std::string& val)
{
    const std::string a = argv[i];
    const std::string ls = l ? std::string("--") + l : std::string(), ss = s ? std::string("-") + s : std::string();
    if ((s && a == ss) || (l && a == ls)) {
        if (i + 1 >= argc) throw std::runtime_error("Flag '" + a + "' requires an argument");
        val = argv[++i];
        return true;
    }
    if (l && a.rfind(ls + "=", 0) == 0) { val = a.substr(ls.size() + 1); return true; }
    if (s && a.rfind(ss, 0) == 0 && a.size() > 2 && a[1] != '-') { val = a.substr(2); return true; }
    return false;
}

}  // namespace

int main(int argc, char** argv)
{
    unsigned threads = 16;
    std::string kmer, reads, ctg, ref, pre, aln, out;
    size_t min_len = 50, eps = 10, cov = 1;
    if (argc <= 1) { usage(); return 0; }
    try {
        for (int i = 1; i < argc; ++i) {
            std::string v;
            const std::string s = argv[i];
            if (s == "-h" || s == "--help") { usage(); return 0; }
            else if (take(argc, argv, i, "t", "thread", v)) threads = (unsigned)std::stoul(v);
            else if (take(argc, argv, i, "k", "kmer", v)) kmer = v;
            else if (take(argc, argv, i, "r", "read", v)) reads = v;
            else if (take(argc, argv, i, "c", "contig", v)) ctg = v;
            else if (take(argc, argv, i, "R", "ref", v)) ref = v;
            else if (take(argc, argv, i, "p", "pre_process", v)) pre = v;
            else if (take(argc, argv, i, "a", "aln", v)) aln = v;
            else if (take(argc, argv, i, "o", "output", v)) out = v;
            else if (take(argc, argv, i, "l", "length", v)) min_len = std::stoul(v);
            else if (take(argc, argv, i, nullptr, "epsilon", v)) eps = std::stoul(v);
            else if (take(argc, argv, i, "v", nullptr, v)) cov = std::stoul(v);
            else throw std::runtime_error("Flag could not be matched: " + s);
        }
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        usage();
        return 1;
    }
    // The GPUs the graph build runs on: all visible devices, or the list AG2_DEVICES names (a device may be
    // named twice: that is how the sharding is tested on a one-GPU box).  One job (its own copy of the inputs, one handle)
    // per device; the reads of a block are sharded contiguously over them, ag2_pg_group_exchange moves the vertex tuples to
    // their owners over NVLink, ag2_pg_group_gather merges the tables into the first job, which traverses.
    std::vector<int> devs;
    if (const char* e = getenv("AG2_DEVICES")) {
        devs = ag2host::parse_device_list(e);
    } else {
        int nd = 0;
        if (ag2_device_count(&nd) == AG2_OK)
            for (int d = 0; d < nd; ++d) devs.push_back(d);
    }
    if (devs.empty()) devs.push_back(0);   // ag2_pg_job_open then reports the missing device
    const int world = (int)devs.size();
    std::vector<ag2_pg_job*> jobs((size_t)world, nullptr);
    std::vector<int> rcs((size_t)world, AG2_OK);
    auto on_all = [&](auto f) {
        if (world == 1) {
            f(0);
            return;
        }
        std::vector<std::thread> th;
        for (int r = 0; r < world; ++r) th.emplace_back([&f, r] { f(r); });
        for (auto& t : th) t.join();
    };
    auto close_all = [&] {
        for (ag2_pg_job* j : jobs) ag2_pg_job_close(j);
    };
    on_all([&](int r) { rcs[(size_t)r] = ag2_pg_job_open(devs[(size_t)r], kmer.c_str(), ctg.c_str(), ref.c_str(), pre.c_str(), aln.c_str(), &jobs[(size_t)r]); });
    int rc = AG2_OK;
    for (int r = 0; r < world; ++r)
        if (rcs[(size_t)r] != AG2_OK) {
            fprintf(stderr, "pagraph (aligngraph2_b200): %s (%d)\n", ag2_pg_job_error(jobs[(size_t)r]), rcs[(size_t)r]);
            close_all();
            return 1;
        }
    ag2_pg_job* job = jobs[0];
    std::vector<ag2_pg*> handles;
    for (ag2_pg_job* j : jobs) handles.push_back(ag2_pg_job_handle(j));
    ag2_pg_params bp;
    ag2_pg_params_default(&bp);
    bp.epsilon = (int64_t)eps;
    bp.cov_filter = (int64_t)cov;
    ag2_pg_travel_params tp;
    ag2_pg_travel_params_default(&tp);
    tp.deviation = (int64_t)eps * 2;
    tp.min_len = (int64_t)min_len;
    tp.threads = (int32_t)threads;
    auto fail = [&](const char* what) {
        fprintf(stderr, "pagraph (aligngraph2_b200): %s failed (%d): %s\n", what, rc, ag2_pg_job_error(job));
        close_all();
        return 1;
    };
    auto fail_pg = [&](const char* what, int r) {
        fprintf(stderr, "pagraph (aligngraph2_b200): %s failed on device %d (%d): %s\n", what, devs[(size_t)r], rcs[(size_t)r], ag2_pg_last_error(handles[(size_t)r]));
        close_all();
        return 1;
    };
    auto first_bad = [&] {
        for (int r = 0; r < world; ++r)
            if (rcs[(size_t)r] != AG2_OK) return r;
        return -1;
    };
    for (int b = 0; b < ag2_pg_job_blocks(job); ++b) {
        std::cout << "Use Ref: " << ag2_pg_job_block_ref(job, b) << std::endl;
        on_all([&](int r) { rcs[(size_t)r] = ag2_pg_job_load_block(jobs[(size_t)r], b, r, world); });
        if (int r = first_bad(); r >= 0) {
            rc = rcs[(size_t)r];
            job = jobs[(size_t)r];
            return fail("ag2_pg_job_load_block");
        }
        if (world == 1) {
            if ((rcs[0] = ag2_pg_build(handles[0], &bp)) != AG2_OK) return fail_pg("ag2_pg_build", 0);
        } else {
            on_all([&](int r) { rcs[(size_t)r] = ag2_pg_extract(handles[(size_t)r], &bp); });
            if (int r = first_bad(); r >= 0) return fail_pg("ag2_pg_extract", r);
            if ((rcs[0] = ag2_pg_group_exchange(handles.data(), world)) != AG2_OK) return fail_pg("ag2_pg_group_exchange", 0);
            on_all([&](int r) { rcs[(size_t)r] = ag2_pg_join(handles[(size_t)r], &bp); });
            if (int r = first_bad(); r >= 0) return fail_pg("ag2_pg_join", r);
            if ((rcs[0] = ag2_pg_group_gather(handles.data(), world)) != AG2_OK) return fail_pg("ag2_pg_group_gather", 0);
        }
        ag2_pg_stats st;
        ag2_pg_get_stats(handles[0], &st);
        std::cout << "\tmerge edge = " << st.edges << "\n\tmerge pos = " << st.positions << std::endl;
        if ((rc = ag2_pg_job_travel(job, b, &tp, out.c_str())) != AG2_OK) return fail("ag2_pg_job_travel");
    }
    if ((rc = ag2_pg_job_write_contig_list(job, out.c_str())) != AG2_OK) return fail("ag2_pg_job_write_contig_list");
    close_all();
    return 0;
}
