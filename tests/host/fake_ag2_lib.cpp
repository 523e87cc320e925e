// fake_ag2_lib.cpp -- TEST DOUBLE of the C ABI (include/ag2_b200.h), for the GPU-less build container only.
//
// Answers the calls the drop-in `mecat2ref` executable makes (ag2_device_count, ag2_ctx_*, ag2_ref_load, ag2_reads_load,
// ag2_index_build, ag2_map_reads, ag2_map_fetch) with the CPU oracle (oracle/ag2_mapper.c), so that the executable's HOST
// logic -- load_fastq batching, the cut of a batch into per-device read ranges, read ids, record order, the thread file,
// result_combine / polish_result -- runs in `pytest -m "not gpu"` against the reference's golden files
// (tests/test_host_flow.py).  FAKE_AG2_DEVICES says how many "devices" ag2_device_count reports.
//
// It is test infrastructure like oracle/ itself: built by the test into a temporary directory, next to a COPY of the
// executable; nothing under aligngraph2_b200/ knows about it, and the product's own library still refuses to run without a
// GPU (tests/test_abi.py, tests/test_host_binary.py::test_without_gpu_exits_1).
#include "../../include/ag2_b200.h"
#include "../../oracle/ag2_oracle.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

struct ag2_ctx {
    int device = 0;
    std::string err, ref, bases;
    std::vector<long> offs;
    orc_index *ix = nullptr;
    std::vector<ag2_record> rec;
    std::string q, s;
};

namespace {
std::mutex g_oracle;   // the oracle is single-threaded test code; the host calls from one thread per device
}

extern "C" {

int ag2_host_alloc(size_t bytes, void **ptr)
{
    *ptr = malloc(bytes ? bytes : 1);
    return *ptr ? AG2_OK : AG2_ENOMEM;
}
void ag2_host_free(void *ptr) { free(ptr); }


int ag2_device_count(int *count)
{
    const char *e = getenv("FAKE_AG2_DEVICES");
    *count = e ? atoi(e) : 1;
    return *count > 0 ? AG2_OK : AG2_ENODEV;
}

int ag2_ctx_create(int device, ag2_ctx **out)
{
    int n = 0;
    if (ag2_device_count(&n) != AG2_OK || device < 0 || device >= n) return AG2_EINVAL;
    *out = new ag2_ctx();
    (*out)->device = device;
    return AG2_OK;
}

void ag2_ctx_destroy(ag2_ctx *ctx)
{
    if (!ctx) return;
    std::lock_guard<std::mutex> lk(g_oracle);
    if (ctx->ix) orc_index_free(ctx->ix);
    delete ctx;
}

const char *ag2_last_error(const ag2_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int ag2_ref_load(ag2_ctx *ctx, const char *ref, int64_t ref_len)
{
    ctx->ref.assign(ref, (size_t)ref_len);
    return AG2_OK;
}

int ag2_reads_load(ag2_ctx *ctx, const char *bases, const int64_t *offs, int64_t n)
{
    if (n <= 0 || offs[0] != 0) return AG2_EINVAL;
    ctx->bases.assign(bases, (size_t)offs[n]);
    ctx->offs.assign(offs, offs + n + 1);
    return AG2_OK;
}

int ag2_index_build(ag2_ctx *ctx, int cbl, double alpha, double beta)
{
    std::lock_guard<std::mutex> lk(g_oracle);
    std::vector<int> rcnt((size_t)1 << 26, 0);
    const long n = (long)ctx->offs.size() - 1;
    const long pre = orc_read_index_prefix(ctx->offs.data(), n);
    orc_read_hist13(ctx->bases.data(), ctx->offs[(size_t)pre], rcnt.data());
    if (ctx->ix) orc_index_free(ctx->ix);
    ctx->ix = orc_index_build(ctx->ref.data(), (long)ctx->ref.size(), rcnt.data(), cbl, alpha, beta);
    return ctx->ix ? AG2_OK : AG2_ENOMEM;
}

int ag2_map_reads(ag2_ctx *ctx, int maxc, int num_output, int64_t *n_records)
{
    if (!ctx->ix) {
        ctx->err = "ag2_map_reads: call ag2_index_build first";
        return AG2_ESTATE;
    }
    std::lock_guard<std::mutex> lk(g_oracle);
    char *text = nullptr;
    size_t text_len = 0;
    FILE *mem = open_memstream(&text, &text_len);
    orc_mapper *m = orc_mapper_new(ctx->ix, maxc, num_output);
    const long n = (long)ctx->offs.size() - 1;
    for (long r = 0; r < n; ++r) {
        const std::string read = ctx->bases.substr((size_t)ctx->offs[(size_t)r], (size_t)(ctx->offs[(size_t)r + 1] - ctx->offs[(size_t)r]));
        orc_map_read(m, (int)r, read.c_str(), (int)read.size(), mem);
    }
    orc_mapper_free(m);
    fclose(mem);
    ctx->rec.clear();
    ctx->q.clear();
    ctx->s.clear();
    // the oracle writes thread-file text: "id\tF|R\tvscore\tqb\tqe\tqs\tsb\tse\n" + the two alignment lines
    const char *p = text, *end = text + text_len;
    while (p < end) {
        const char *l1 = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *l2 = l1 ? (const char *)memchr(l1 + 1, '\n', (size_t)(end - l1 - 1)) : nullptr;
        const char *l3 = l2 ? (const char *)memchr(l2 + 1, '\n', (size_t)(end - l2 - 1)) : nullptr;
        if (!l3) break;
        ag2_record r;
        memset(&r, 0, sizeof r);
        char dir = 'F';
        long sb = 0, se = 0;
        if (sscanf(p, "%d\t%c\t%d\t%d\t%d\t%d\t%ld\t%ld", &r.read, &dir, &r.vscore, &r.qb, &r.qe, &r.qs, &sb, &se) != 8) break;
        r.ok = 1;
        r.strand = dir == 'R';
        r.sb = sb;
        r.se = se;
        r.aln_len = (int32_t)(l2 - l1 - 1);
        r.aln_off = (int64_t)ctx->q.size();
        ctx->q.append(l1 + 1, (size_t)r.aln_len);
        ctx->s.append(l2 + 1, (size_t)(l3 - l2 - 1));
        ctx->rec.push_back(r);
        p = l3 + 1;
    }
    free(text);
    *n_records = (int64_t)ctx->rec.size();
    return AG2_OK;
}

int ag2_map_fetch(ag2_ctx *ctx, ag2_record *rec_out, char *qaln_out, char *saln_out, int64_t aln_cap, int64_t *aln_used)
{
    memcpy(rec_out, ctx->rec.data(), ctx->rec.size() * sizeof(ag2_record));
    if (aln_used) *aln_used = (int64_t)ctx->q.size();
    if (qaln_out && saln_out) {
        if (aln_cap < (int64_t)ctx->q.size()) return AG2_ECAP;
        memcpy(qaln_out, ctx->q.data(), ctx->q.size());
        memcpy(saln_out, ctx->s.data(), ctx->s.size());
    }
    return AG2_OK;
}

} // extern "C"
