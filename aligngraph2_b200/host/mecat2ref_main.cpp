// mecat2ref_main.cpp -- drop-in replacement of the mecat2ref+ executable: same argv, same files,
// same exit codes; the per-read work runs on the GPU through the C ABI of include/ag2_b200.h.
//
// Mirrors mecat_plus/MECAT-master_1/src/mecat2ref/mecat2ref.cpp:
//   param_read_t :98-219 (getopt string, checks, -b clamped to -n, working dir created), firsttask :398-434
//   (reads -> <wrk>/0.fq, reference -> <wrk>/ref.fq, ./config.txt), main :951-1016, result_combine :523-599,
//   polish_result :625-867, get_chr_id :444-462, output_query_results :499-520; and of
//   mecat2ref_impl_large.cpp: creat_ref_index's FASTA reading and chrindex.txt :421-450, load_fastq's batching
//   :1965-1991, the timing lines of config.txt :2017-2033,2133-2138.
//
// Differences that are deliberate and visible:
//   * -t is accepted and recorded, but the mapping always behaves like `-t 1`: all records go to <wrk>/1.r in read
//     order (2.r..N.r and refN.r are created empty).  The reference's own per-read results do not depend on -t
//     (SURVEY F6), only their order and, through polish_result's "last group of every thread file is not
//     filtered" rule (:854), which groups escape the filter -- with one thread file that is deterministic.
//   * the reads of every load_fastq batch are sharded over all visible GPUs (AG2_DEVICES names another set), one host
//     thread per GPU, records written in device order = read order: the files do not depend on the number of GPUs.
//   * the next load_fastq batch is parsed on a second thread while the current one is mapped and printed, and
//     result_combine / polish_result run side by side (separate output files): same bytes, less waiting on text.
//   * -x is parsed and ignored, as in the reference.
//   * no CPU fallback: without a usable GPU the program exits 1 (AlignGraph2.py:280-296 then falls back to the
//     vanilla mecat2ref exactly as it does for any mecat2ref+ failure).
// AG2_SKIP_MAP=1 in the environment skips the GPU stage and reuses the thread files already in <wrk> (used by
// the CPU-side tests of the file handling, result_combine and polish_result).
#include "../../include/ag2_b200.h"
#include "shard_split.h"

#include <dirent.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cstdarg>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int kSVM = 100000;              // reads per batch (mecat2ref_defs.h:27)
constexpr long kMAXSTR = 1000000000L;     // characters per batch (:25)

struct Options {
    const char *reads = nullptr, *reference = nullptr, *wrk_dir = nullptr, *output = nullptr, *refoutput = nullptr;
    int num_cores = 1, num_candidates = 10, num_output = 10, output_format = 0, tech = 0, block = 200;
    double alpha = 0.5, beta = 2.0, delta = 0.9;
};

const char *prog_name = "mecat2ref";

void print_usage()
{
    fprintf(stderr, "\n\nusage:\n%s [-d reads] [-r reference] [-o output][-p refoutput] [-w working dir] [-t threads]\n\n", prog_name);
    fprintf(stderr, "options:\n-d <string>\treads file name\n-r <string>\treference file name\n-o <string>\toutput file name\n");
    fprintf(stderr, "-p <string>\trefoutput file name\n-w <string>\tworking folder name, will be created if not exist\n");
    fprintf(stderr, "-t <integer>\tnumber of cput threads\n\t\tdefault: 1\n-n <integer>\tnumber of of candidates for gap extension\n\t\tdefault: 10\n");
    fprintf(stderr, "-b <integer>\toutput the best b alignments\n\t\tdefault: 10\n-m <0/1/2>\toutput format: 0 = ref, 1 = m4, 2 = sam\n\t\tdefault: 0\n");
    fprintf(stderr, "-x <0/1>\tsequencing technology: 0 = pacbio, 1 = nanopore\n\t\tdefault: 0\n");
    fprintf(stderr, "-l <real>\tlower bound of k-mer scoring function for mecat2ref+\n\t\tdefault: 0.5\n");
    fprintf(stderr, "-u <real>\tupper bound of k-mer scoring function for mecat2ref+\n\t\tdefault: 2.0\n");
    fprintf(stderr, "-z <integer>\tsize of similar genome blocks for mecat2ref+\n\t\tdefault: 200\n-y <integer>\tthreshold for alignment scoring\n\t\tdefault: 0.9\n");
}

int parse_options(int argc, char **argv, Options &o)
{
    opterr = 0;
    int c;
    while ((c = getopt(argc, argv, "d:r:w:o:p:t:n:b:m:x:l:u:z:y:")) != -1) {
        switch (c) {
        case 'd': o.reads = optarg; break;
        case 'r': o.reference = optarg; break;
        case 'w': o.wrk_dir = optarg; break;
        case 'o': o.output = optarg; break;
        case 'p': o.refoutput = optarg; break;
        case 't': o.num_cores = atoi(optarg); break;
        case 'n': o.num_candidates = atoi(optarg); break;
        case 'b': o.num_output = atoi(optarg); break;
        case 'm': o.output_format = atoi(optarg); break;
        case 'x':
            if (optarg[0] == '0') o.tech = 0;
            else if (optarg[0] == '1') o.tech = 1;
            else { fprintf(stderr, "Invalid argument to option 'x': %s\n", optarg); abort(); }
            break;
        case 'l': o.alpha = atof(optarg); break;
        case 'u': o.beta = atof(optarg); break;
        case 'z': o.block = atoi(optarg); break;
        case 'y': o.delta = atof(optarg); break;
        case ':': fprintf(stderr, "Error: unrecogised option '%c'\n", (char)optopt); return -1;
        case '?': fprintf(stderr, "Error: argument to option '%c' is missing!\n", (char)optopt); return -1;
        }
    }
    const char *msg = nullptr;
    if (!o.reads) msg = "dataset must be specified";
    else if (!o.reference) msg = "reference must be specified";
    else if (!o.output) msg = "output must be specified";
    else if (!o.refoutput) msg = "refoutput must be specified";
    else if (!o.wrk_dir) msg = "working directory must be specified";
    else if (o.num_cores < 1) msg = "cpu cores must be > 0";
    else if (o.num_candidates < 1) msg = "candidates must be > 0";
    else if (o.num_output < 1) msg = "output alignments must be > 0";
    else if (o.num_candidates > 16) msg = "candidates must be <= 16";
    if (msg) {
        fprintf(stderr, "Error: %s\n", msg);
        return -1;
    }
    if (o.num_output > o.num_candidates) {
        fprintf(stderr, "warning: number of output (%d) is greater than number of candidates (%d), we reset it to %d", o.num_output,
                o.num_candidates, o.num_candidates);
        o.num_output = o.num_candidates;
    }
    DIR *d = opendir(o.wrk_dir);
    if (!d) {
        if (mkdir(o.wrk_dir, S_IRWXU) == -1) {
            fprintf(stderr, "Fail to create folder %s!\n", o.wrk_dir);
            return -1;
        }
    } else {
        closedir(d);
    }
    return 1;
}

// chang_fastqfile / change_ref_fq (mecat2ref.cpp:272-395): FASTA -> ids 0.., FASTQ (4 lines) -> ids 1..
// `sink` (may be empty) sees every read as it is written: (id, sequence, length).
int convert_to_fq(const char *in_path, const std::string &out_path, const std::function<void(int, const char *, size_t)> &sink = nullptr)
{
    FILE *fp = fopen(in_path, "r");
    if (!fp) { fprintf(stderr, "failed to open file %s for reading.\n", in_path); exit(1); }
    FILE *ot = fopen(out_path.c_str(), "w");
    if (!ot) { fprintf(stderr, "failed to open file %s for writing.\n", out_path.c_str()); exit(1); }
    std::vector<char> obuf(1 << 24);
    setvbuf(ot, obuf.data(), _IOFBF, obuf.size());
    int kk = 0;
    int ch = getc_unlocked(fp);   // one thread per FILE here: the unlocked getc is 3x the locked one on a 250 Mb reference
    std::string one;
    if (ch == '>') {
        for (; ch != EOF; ch = getc_unlocked(fp)) {
            if (ch == '>') {
                while ((ch = getc_unlocked(fp)) != EOF && ch != '\n') {}
                if (ch == '\n') ungetc(ch, fp);
                if (kk > 0) {
                    fprintf(ot, "%d\t%d\t%s\n", kk - 1, (int)one.size(), one.c_str());
                    if (sink) sink(kk - 1, one.data(), one.size());
                }
                one.clear();
                kk++;
            } else if (ch != '\n' && ch != '\r') {
                one.push_back((char)ch);
            }
        }
        fprintf(ot, "%d\t%d\t%s\n", kk - 1, (int)one.size(), one.c_str());
        if (sink) sink(kk - 1, one.data(), one.size());
    } else {
        fseek(fp, 0L, SEEK_SET);
        // The reference reads a record with fscanf("%[^\n]s") / fscanf("%s\n") pairs (mecat2ref.cpp:317): the name and '+'
        // lines up to the newline, the sequence and the quality as ONE whitespace-delimited token each, all whitespace after
        // a token (blank lines included) swallowed.  Lines are read whole (getline) while a record is "clean" -- then both
        // readings agree -- and from the first record that is not (leading / inner blanks, blank lines) the rest of the file
        // goes through scan_line / scan_token, which are those two conversions character by character.
        auto is_ws = [](int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; };
        auto scan_line = [&]() -> int {           // "%[^\n]": 1 = matched, 0 = the next character is a newline (left unread), EOF
            int c = getc_unlocked(fp);
            if (c == EOF) return EOF;
            if (c == '\n') {
                ungetc(c, fp);
                return 0;
            }
            while (c != EOF && c != '\n') c = getc_unlocked(fp);
            if (c == '\n') ungetc(c, fp);
            return 1;
        };
        auto scan_token = [&](std::string &out) -> int {   // "%s\n"
            int c;
            do c = getc_unlocked(fp);
            while (c != EOF && is_ws(c));
            if (c == EOF) return EOF;
            out.clear();
            while (c != EOF && !is_ws(c)) {
                out.push_back((char)c);
                c = getc_unlocked(fp);
            }
            while (c != EOF && is_ws(c)) c = getc_unlocked(fp);
            if (c != EOF) ungetc(c, fp);
            return 1;
        };
        // token + trailing blanks + newline, nothing else; *len = token length
        auto clean_token_line = [&](const char *l, ssize_t n, size_t *len) {
            ssize_t p = 0;
            while (p < n && !is_ws((unsigned char)l[p])) ++p;
            *len = (size_t)p;
            if (p == 0) return false;
            for (ssize_t q = p; q < n; ++q)
                if (!is_ws((unsigned char)l[q])) return false;
            return n > 0 && l[n - 1] == '\n';
        };
        char *l1 = nullptr, *l2 = nullptr, *l3 = nullptr, *l4 = nullptr;
        size_t c1 = 0, c2 = 0, c3 = 0, c4 = 0;
        bool clean = true;
        while (clean) {
            const long at = ftell(fp);
            const ssize_t n1 = getline(&l1, &c1, fp), n2 = n1 < 0 ? -1 : getline(&l2, &c2, fp), n3 = n2 < 0 ? -1 : getline(&l3, &c3, fp),
                          n4 = n3 < 0 ? -1 : getline(&l4, &c4, fp);
            size_t slen = 0, qlen = 0;
            bool ok = n4 >= 0 && n1 > 0 && l1[n1 - 1] == '\n' && clean_token_line(l2, n2, &slen) && n3 > 0 && !is_ws((unsigned char)l3[0]) &&
                      l3[n3 - 1] == '\n' && clean_token_line(l4, n4, &qlen);
            if (ok) {   // the token's trailing skip must stop at the first character of the next record
                const int c = getc_unlocked(fp);
                if (c != EOF) {
                    ungetc(c, fp);
                    if (is_ws(c)) ok = false;
                }
            }
            if (!ok) {
                fseek(fp, at, SEEK_SET);
                clean = false;
                break;
            }
            fprintf(ot, "%d\t%d\t", ++kk, (int)slen);
            fwrite(l2, 1, slen, ot);
            fputc('\n', ot);
            if (sink) sink(kk, l2, slen);
        }
        free(l1);
        free(l2);
        free(l3);
        free(l4);
        std::string seq, qual;
        while (scan_line() != EOF && scan_token(seq) != EOF && scan_line() != EOF && scan_token(qual) != EOF) {
            fprintf(ot, "%d\t%d\t%s\n", ++kk, (int)seq.size(), seq.c_str());
            if (sink) sink(kk, seq.data(), seq.size());
        }
    }
    fclose(fp);
    fclose(ot);
    return kk;
}

struct ChrInfo {
    long start, size;
    char name[64];
};

// creat_ref_index's FASTA reading (impl_large.cpp:421-450): concatenated sequence, upper-cased above 'Z',
// <wrk>/chrindex.txt
void load_reference(const char *path, const std::string &wrk, std::string &seq)
{
    FILE *fasta = fopen(path, "r");
    if (!fasta) { fprintf(stderr, "failed to open file %s for reading.\n", path); exit(1); }
    FILE *idx = fopen((wrk + "/chrindex.txt").c_str(), "w");
    if (!idx) { fprintf(stderr, "failed to open %s/chrindex.txt for writing.\n", wrk.c_str()); exit(1); }
    long rsize = 0, count = 0;
    seq.clear();
    {
        struct stat st;
        if (fstat(fileno(fasta), &st) == 0 && st.st_size > 0) seq.reserve((size_t)st.st_size);
    }
    char nameall[1 << 16];
    for (int ch = getc_unlocked(fasta); ch != EOF; ch = getc_unlocked(fasta)) {
        if (ch == '>') {
            if (fscanf(fasta, "%65535[^\n]", nameall) != 1) nameall[0] = 0;
            if (rsize) fprintf(idx, "%ld\n", rsize);
            rsize = 0;
            size_t i;
            for (i = 0; i < strlen(nameall); i++)
                if (nameall[i] == ' ' || nameall[i] == '\t') break;
            nameall[i] = '\0';
            fprintf(idx, "%ld\t%s\t", count, nameall);
        } else if (ch != '\n' && ch != '\r') {
            if (ch > 'Z') ch = toupper(ch);
            seq.push_back((char)ch);
            ++count;
            ++rsize;
        }
    }
    fclose(fasta);
    fprintf(idx, "%ld\n", rsize);
    fprintf(idx, "%ld\t%s\n", count, "FileEnd");
    fclose(idx);
}

std::vector<ChrInfo> read_chrindex(const std::string &wrk)
{
    const std::string path = wrk + "/chrindex.txt";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) { fprintf(stderr, "failed to open file %s for reading.\n", path.c_str()); abort(); }
    char buffer[1024];
    int num_chr = 0;
    while (fgets(buffer, 1024, f)) ++num_chr;
    --num_chr;
    fseek(f, 0L, SEEK_SET);
    std::vector<ChrInfo> v((size_t)(num_chr > 0 ? num_chr : 0));
    for (int i = 0; i < num_chr; ++i) {
        const int flag = fscanf(f, "%ld\t%63s\t%ld\n", &v[i].start, v[i].name, &v[i].size);
        assert(flag == 3);
        (void)flag;
    }
    fclose(f);
    return v;
}

int get_chr_id(const std::vector<ChrInfo> &chr, long offset) // mecat2ref.cpp:444-462
{
    const int num_chr = (int)chr.size();
    int left = 0, right = num_chr, mid = 0;
    while (left < right) {
        mid = (left + right) >> 1;
        if (offset >= chr[mid].start) {
            if (mid == num_chr - 1) break;
            if (offset < chr[mid + 1].start) break;
            left = mid + 1;
        } else {
            right = mid;
        }
    }
    return mid;
}

struct TempResult {
    int read_id = 0, vscore = 0, qb = 0, qe = 0, qs = 0;
    char read_dir = 0;
    long sb = 0, se = 0;
    std::string qmap, smap;
};

bool load_temp_result(TempResult &r, FILE *in, char *&line, size_t &cap) // output.cpp:270-289
{
    if (getline(&line, &cap, in) < 0) return false;
    if (sscanf(line, "%d\t%c\t%d\t%d\t%d\t%d\t%ld\t%ld", &r.read_id, &r.read_dir, &r.vscore, &r.qb, &r.qe, &r.qs, &r.sb, &r.se) != 8) return false;
    ssize_t n = getline(&line, &cap, in);
    if (n < 0) return false;
    r.qmap.assign(line, (size_t)n);
    while (!r.qmap.empty() && r.qmap.back() == '\n') r.qmap.pop_back();
    n = getline(&line, &cap, in);
    if (n < 0) return false;
    r.smap.assign(line, (size_t)n);
    while (!r.smap.empty() && r.smap.back() == '\n') r.smap.pop_back();
    return true;
}

// One record as the writers see it: TempResult without owning its strings (they stay in the buffers the GPU results
// came home in, or in the TempResult a thread file line was parsed into).
struct RecView {
    int read_id = 0, vscore = 0, qb = 0, qe = 0, qs = 0;
    char read_dir = 0;
    long sb = 0, se = 0;
    const char *qmap = nullptr, *smap = nullptr;
    int len = 0;
};

RecView view_of(const TempResult &r)
{
    RecView v;
    v.read_id = r.read_id, v.vscore = r.vscore, v.qb = r.qb, v.qe = r.qe, v.qs = r.qs;
    v.read_dir = r.read_dir;
    v.sb = r.sb, v.se = r.se;
    v.qmap = r.qmap.data(), v.smap = r.smap.data();
    v.len = (int)r.qmap.size();
    return v;
}

void append_fmt(std::string &out, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void append_fmt(std::string &out, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    const int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    out.append(buf, (size_t)std::min<int>(n, (int)sizeof buf - 1));
}

void append_cigar(const RecView &r, std::string &out)
{
    if (r.qb) append_fmt(out, "%dH", r.qb);
    int i = 0, j;
    const int n = r.len;
    while (i < n) {
        if (r.qmap[i] == '-') {
            j = i + 1;
            while (j < n && r.qmap[j] == '-') ++j;
            append_fmt(out, "%dD", j - i);
        } else if (r.smap[i] == '-') {
            j = i + 1;
            while (j < n && r.smap[j] == '-') ++j;
            append_fmt(out, "%dI", j - i);
        } else {
            j = i + 1;
            while (j < n && r.qmap[j] != '-' && r.smap[j] != '-') ++j;
            append_fmt(out, "%dM", j - i);
        }
        i = j;
    }
    if (r.qe != r.qs) append_fmt(out, "%dH", r.qs - r.qe);
}

// output_one_result (output.cpp:6-43 ref, :45-88 m4, :150-186 sam), appended to a buffer
void format_one(const RecView &r, const ChrInfo &c, int format, std::string &out)
{
    const long sstart = r.sb - c.start, send = r.se - c.start;
    int qb = r.qb, qe = r.qe;
    if (format == 0) {
        if (r.read_dir == 'R') {
            qb = r.qs - r.qe;
            qe = r.qs - r.qb;
        }
        append_fmt(out, "%d\t%s\t%c\t%d\t%d\t%d\t%d\t%ld\t%ld\t%ld\n", r.read_id, c.name, r.read_dir == 'R' ? 'R' : 'F', r.vscore, qb, qe, r.qs,
                   sstart, send, c.size);
        out.append(r.qmap, (size_t)r.len);
        out.push_back('\n');
        out.append(r.smap, (size_t)r.len);
        out.push_back('\n');
    } else if (format == 1) {
        if (r.read_dir == 'R') {
            qb = r.qs - r.qe;
            qe = r.qs - r.qb;
        }
        double ident = 0.0;
        const int n = r.len;
        for (int i = 0; i < n; ++i)
            if (r.qmap[i] == r.smap[i]) ident += 1.0;
        ident = ident / n;
        ident *= 100.0;
        append_fmt(out, "%d\t%s\t%.4f\t%d\t%d\t%d\t%d\t%d\t0\t%ld\t%ld\t%ld\n", r.read_id, c.name, ident, r.vscore, r.read_dir == 'F' ? 0 : 1, qb,
                   qe, r.qs, sstart, send, c.size);
    } else if (format == 2) {
        append_fmt(out, "%d\t%d\t%s\t%ld\t255\t", r.read_id, r.read_dir == 'R' ? 0x10 : 0, c.name, sstart + 1);
        append_cigar(r, out);
        out.append("\t*\t0\t0\t");
        for (int i = 0; i < r.len; ++i)
            if (r.qmap[i] != '-') out.push_back(r.qmap[i]);
        out.append("\t*\n");
    }
}

// output_temp_result (output.cpp:237-251): the `<wrk>/N.r` text of one record
void format_temp(const RecView &r, std::string &out)
{
    append_fmt(out, "%d\t%c\t%d\t%d\t%d\t%d\t%ld\t%ld\n", r.read_id, r.read_dir, r.vscore, r.qb, r.qe, r.qs, r.sb, r.se);
    out.append(r.qmap, (size_t)r.len);
    out.push_back('\n');
    out.append(r.smap, (size_t)r.len);
    out.push_back('\n');
}

void sam_header(const std::vector<ChrInfo> &chr, int argc, char **argv, FILE *out)
{
    fprintf(out, "@HD\tVN:1.4\tSO:unknown\tGO:query\n");
    for (const ChrInfo &c : chr) fprintf(out, "@SQ\tSN:%s\tLN:%ld\n", c.name, c.size);
    fprintf(out, "@PG\tID:0\tVN:0.0.1\tCL:");
    for (int i = 0; i < argc; ++i) fprintf(out, "%s ", argv[i]);
    fprintf(out, "\tPN:mecat2ref\n");
}

void format_query_results(const std::vector<ChrInfo> &chr, const std::vector<const RecView *> &p, int num_output, int format, std::string &out)
{
    int cnt = 0;
    for (const RecView *r : p) { // mecat2ref.cpp:499-520
        format_one(*r, chr[(size_t)get_chr_id(chr, r->sb)], format, out);
        if (++cnt == num_output) break;
    }
}

// The co-linearity vote of polish_result over one read group (mecat2ref.cpp:746-848): which records are written, in order.
void polish_group(const std::vector<RecView> &pptr, const std::vector<ChrInfo> &chr, double delta, std::vector<const RecView *> &outp)
{
    int vote[16] = {0}, mark[16] = {0};
    int flag3 = 0, flag4 = 0;
    const int num_results = (int)pptr.size();
    outp.clear();
    for (int i = 0; i < num_results; i++)
        for (int j = i + 1; j < num_results; j++) {
            const int sid = get_chr_id(chr, pptr[i].sb), sid2 = get_chr_id(chr, pptr[j].sb);
            if (sid == sid2 && labs(pptr[j].qb - pptr[i].qb) > 1000 && labs(pptr[i].qe - pptr[j].qe) > 1000 && pptr[i].sb != pptr[j].sb &&
                fabs((double)((pptr[i].qb - pptr[j].qb) / (pptr[i].sb - pptr[j].sb) - 1)) < delta) {
                vote[i]++;
                vote[j]++;
                mark[i] = 1;
                mark[j] = 1;
            }
        }
    for (int k = 0; k < num_results; k++) {
        int delete_flag = 0;
        if (mark[k] == 0) outp.push_back(&pptr[k]);
        if (mark[k] == 1) {
            int maxi_vote = vote[k], maxi = k;
            const int sid = get_chr_id(chr, pptr[k].sb);
            for (int p = k + 1; p < num_results; p++) {
                if (sid != get_chr_id(chr, pptr[p].sb)) continue;
                const int lk = pptr[k].qe - pptr[k].qb, lp = pptr[p].qe - pptr[p].qb;
                if (labs(pptr[p].qb - pptr[k].qb) < 1500) {
                    mark[p] = 2;
                    delete_flag = 1;
                    if (labs(lk - lp) > 3000) {
                        maxi = lk > lp ? k : p;
                    } else {
                        if (maxi_vote > vote[p]) maxi = k;
                        if (maxi_vote == vote[p]) maxi = lk > lp ? k : p;
                        if (maxi_vote < vote[p]) {
                            maxi = p;
                            maxi_vote = vote[p];
                        }
                    }
                } else {
                    for (int q = p + 1; q < num_results; q++)
                        if (sid == get_chr_id(chr, pptr[q].sb)) {
                            flag3 = 1;
                            break;
                        }
                    if (flag3 == 0) outp.push_back(&pptr[k]);
                    flag3 = 0;
                }
            }
            if (delete_flag == 1) outp.push_back(&pptr[maxi]);
            for (int w = k + 1; w < num_results; w++)
                if (sid == get_chr_id(chr, pptr[w].sb)) {
                    flag4 = 1;
                    break;
                }
            if (flag4 == 0) outp.push_back(&pptr[k]);
            flag4 = 0;
        }
    }
}

void write_all(FILE *f, const std::string &buf)
{
    if (!buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
        fprintf(stderr, "mecat2ref (aligngraph2_b200): write failed\n");
        abort();
    }
}

// result_combine (mecat2ref.cpp:523-599): thread files -> -o, grouped by read id, first num_output of a group.
// File-based form: used when the thread files are all there is (AG2_SKIP_MAP); a mapping run writes -o / -p from the
// records in memory (ResultWriter below), through the same format_* / polish_group functions.
void result_combine(const Options &o, const std::vector<ChrInfo> &chr, int argc, char **argv)
{
    fprintf(stderr, "output file name: %s\n", o.output);
    FILE *out = fopen(o.output, "w");
    if (!out) { fprintf(stderr, "failed to open file %s for writing.\n", o.output); abort(); }
    if (o.output_format == 2) sam_header(chr, argc, argv, out);
    char *line = nullptr;
    size_t cap = 0;
    std::string buf;
    for (int i = 1; i <= o.num_cores; ++i) {
        const std::string path = std::string(o.wrk_dir) + "/" + std::to_string(i) + ".r";
        FILE *f = fopen(path.c_str(), "r");
        if (!f) { fprintf(stderr, "failed to open file %s for reading.\n", path.c_str()); abort(); }
        std::vector<TempResult> group;
        TempResult t;
        bool ok = load_temp_result(t, f, line, cap);
        int last = 0;
        if (ok) {
            group.push_back(t);
            last = t.read_id;
        }
        auto flush = [&]() {
            std::vector<RecView> v;
            for (const TempResult &g : group) v.push_back(view_of(g));
            std::vector<const RecView *> p;
            for (const RecView &g : v) p.push_back(&g);
            buf.clear();
            format_query_results(chr, p, o.num_output, o.output_format, buf);
            write_all(out, buf);
            group.clear();
        };
        while (ok) {
            ok = load_temp_result(t, f, line, cap);
            if (!ok) break;
            if (t.read_id != last) flush();
            last = t.read_id;
            group.push_back(t);
        }
        if (!group.empty()) flush();
        fclose(f);
    }
    free(line);
    fclose(out);
}

// polish_result (mecat2ref.cpp:625-867): thread files -> -p with the co-linearity vote per read group.  The last
// group of every thread file is written unfiltered (:854), as in the reference.
void polish_result(const Options &o, const std::vector<ChrInfo> &chr, int argc, char **argv)
{
    fprintf(stderr, "output file name: %s\n", o.refoutput);
    FILE *out = fopen(o.refoutput, "w");
    if (!out) { fprintf(stderr, "failed to open file %s for writing.\n", o.refoutput); abort(); }
    if (o.output_format == 2) sam_header(chr, argc, argv, out);
    char *line = nullptr;
    size_t cap = 0;
    std::string buf;
    for (int ww = 1; ww <= o.num_cores; ww++) {
        const std::string path = std::string(o.wrk_dir) + "/" + std::to_string(ww) + ".r";
        FILE *f = fopen(path.c_str(), "r");
        if (!f) { fprintf(stderr, "failed to open file %s for reading\n", path.c_str()); abort(); }
        std::vector<TempResult> pptr;
        TempResult t;
        bool rok = load_temp_result(t, f, line, cap);
        int last_id = 0;
        if (rok) {
            pptr.push_back(t);
            last_id = t.read_id;
        }
        auto emit = [&](bool filtered) {
            std::vector<RecView> v;
            for (const TempResult &g : pptr) v.push_back(view_of(g));
            std::vector<const RecView *> outp;
            if (filtered) polish_group(v, chr, o.delta, outp);
            else
                for (const RecView &g : v) outp.push_back(&g);
            buf.clear();
            format_query_results(chr, outp, o.num_output, o.output_format, buf);
            write_all(out, buf);
            pptr.clear();
        };
        while (rok) {
            rok = load_temp_result(t, f, line, cap);
            if (!rok) break;
            if (t.read_id != last_id) emit(true);
            last_id = t.read_id;
            pptr.push_back(t);
        }
        if (!pptr.empty()) emit(false);
        fclose(f);
    }
    free(line);
    fclose(out);
}

// The three result files of a mapping run, written from the records in memory: <wrk>/1.r as
// output_temp_result prints it, -o as result_combine and -p as polish_result would make them from that file -- without
// writing it, reading it back twice and parsing 2 x 10 kB of text per record.  A batch is cut at read-group boundaries into
// pieces that worker threads format side by side; the three files are then written by three threads.  polish_result's
// "last group of the thread file is not filtered" (:854) is kept by holding the last group of every batch back until it is
// known whether another batch follows.
class ResultWriter {
public:
    ResultWriter(const Options &o, const std::vector<ChrInfo> &chr, int argc, char **argv) : o_(o), chr_(chr)
    {
        const std::string wrk = o.wrk_dir;
        f_r_ = open_or_die((wrk + "/1.r").c_str());
        same_ = strcmp(o.output, o.refoutput) == 0;   // the reference writes -o, then -p over it
        fprintf(stderr, "output file name: %s\n", o.output);
        if (!same_) f_o_ = open_or_die(o.output);
        fprintf(stderr, "output file name: %s\n", o.refoutput);
        f_p_ = open_or_die(o.refoutput);
        if (o.output_format == 2) {
            if (f_o_) sam_header(chr, argc, argv, f_o_);
            sam_header(chr, argc, argv, f_p_);
        }
        workers_ = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        if (const char *e = getenv("AG2_WRITER_THREADS")) workers_ = std::max(1, atoi(e));
    }

    // records of one batch, in thread-file order; the views point into buffers that outlive the call
    void add_batch(const std::vector<RecView> &recs)
    {
        // read groups: [g[k], g[k+1])
        std::vector<size_t> g;
        for (size_t i = 0; i < recs.size(); ++i)
            if (i == 0 || recs[i].read_id != recs[i - 1].read_id) g.push_back(i);
        g.push_back(recs.size());
        const size_t ngroups = g.size() - 1;
        if (ngroups == 0) return;
        // -p: the carried group of the previous batch is now known not to be the last one of the file
        std::string carried_p;
        if (!carry_.empty()) {
            format_polished(carry_, true, carried_p);
            carry_.clear();
            carry_store_.clear();
        }
        const size_t per = std::max<size_t>(1, (ngroups + (size_t)workers_ * 4 - 1) / ((size_t)workers_ * 4));
        const size_t npieces = (ngroups + per - 1) / per;
        std::vector<std::string> tr(npieces), to(npieces), tp(npieces);
        std::vector<std::thread> th;
        std::atomic<size_t> next{0};
        auto work = [&]() {
            std::vector<RecView> grp;
            std::vector<const RecView *> sel;
            for (;;) {
                const size_t pc = next.fetch_add(1);
                if (pc >= npieces) break;
                const size_t g0 = pc * per, g1 = std::min(ngroups, g0 + per);
                size_t bytes = 0;
                for (size_t i = g[g0]; i < g[g1]; ++i) bytes += 2 * (size_t)recs[i].len + 96;
                tr[pc].reserve(bytes);
                if (f_o_) to[pc].reserve(bytes + 64 * (g[g1] - g[g0]));
                tp[pc].reserve(bytes + 64 * (g[g1] - g[g0]));
                for (size_t k = g0; k < g1; ++k) {
                    grp.assign(recs.begin() + (long)g[k], recs.begin() + (long)g[k + 1]);
                    for (const RecView &r : grp) format_temp(r, tr[pc]);
                    if (f_o_) {
                        sel.clear();
                        for (const RecView &r : grp) sel.push_back(&r);
                        format_query_results(chr_, sel, o_.num_output, o_.output_format, to[pc]);
                    }
                    if (k + 1 < ngroups) {   // the batch's last group waits for the next batch (or the end)
                        polish_group(grp, chr_, o_.delta, sel);
                        format_query_results(chr_, sel, o_.num_output, o_.output_format, tp[pc]);
                    }
                }
            }
        };
        for (int w = 1; w < workers_; ++w) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        // hold the last group back: copy it (its buffers go away with the batch)
        for (size_t i = g[ngroups - 1]; i < g[ngroups]; ++i) {
            carry_store_.emplace_back(std::string(recs[i].qmap, (size_t)recs[i].len), std::string(recs[i].smap, (size_t)recs[i].len));
            carry_.push_back(recs[i]);
        }
        for (size_t i = 0; i < carry_.size(); ++i) {
            carry_[i].qmap = carry_store_[i].first.data();
            carry_[i].smap = carry_store_[i].second.data();
        }
        // three files, three writers; the pieces are megabytes each: straight write(2), no stdio copy
        std::thread wr([&] { for (const std::string &b : tr) write_fd(f_r_, b); });
        std::thread wo([&] { if (f_o_) for (const std::string &b : to) write_fd(f_o_, b); });
        write_fd(f_p_, carried_p);
        for (const std::string &b : tp) write_fd(f_p_, b);
        wr.join();
        wo.join();
    }

    void finish()
    {
        if (!carry_.empty()) {   // the last group of the file: written as it is (mecat2ref.cpp:854)
            std::string p;
            format_polished(carry_, false, p);
            write_fd(f_p_, p);
            carry_.clear();
        }
        fclose(f_r_);
        if (f_o_) fclose(f_o_);
        fclose(f_p_);
        f_r_ = f_o_ = f_p_ = nullptr;
    }

private:
    static FILE *open_or_die(const char *path)
    {
        FILE *f = fopen(path, "w");
        if (!f) { fprintf(stderr, "failed to open file %s for writing.\n", path); abort(); }
        setvbuf(f, nullptr, _IOFBF, 1 << 22);
        return f;
    }
    static void write_fd(FILE *f, const std::string &buf)
    {
        fflush(f);   // whatever stdio still holds (the SAM header) goes first
        const int fd = fileno(f);
        for (size_t done = 0; done < buf.size();) {
            const ssize_t n = write(fd, buf.data() + done, buf.size() - done);
            if (n < 0) {
                fprintf(stderr, "mecat2ref (aligngraph2_b200): write failed\n");
                abort();
            }
            done += (size_t)n;
        }
    }
    void format_polished(const std::vector<RecView> &grp, bool filtered, std::string &out)
    {
        std::vector<const RecView *> sel;
        if (filtered) polish_group(grp, chr_, o_.delta, sel);
        else
            for (const RecView &r : grp) sel.push_back(&r);
        format_query_results(chr_, sel, o_.num_output, o_.output_format, out);
    }
    const Options &o_;
    const std::vector<ChrInfo> &chr_;
    FILE *f_r_ = nullptr, *f_o_ = nullptr, *f_p_ = nullptr;
    bool same_ = false;
    int workers_ = 1;
    std::vector<RecView> carry_;
    std::vector<std::pair<std::string, std::string>> carry_store_;
};

double now_sec()
{
    struct timeval t;
    gettimeofday(&t, nullptr);
    return t.tv_sec + t.tv_usec * 1e-6;
}

void die_ag2(ag2_ctx *ctx, const char *what, int rc)
{
    fprintf(stderr, "mecat2ref (aligngraph2_b200): %s failed (%d): %s\n", what, rc, ag2_last_error(ctx));
    fflush(stderr);
    _exit(1);   // not exit(): the read-ahead thread may be inside getline on a FILE that exit()'s stdio teardown would touch
}

// The GPUs the mapping runs on (reads shard across the GPUs of one box, no data-path collective): all visible
// devices, or the list AG2_DEVICES names ("0,1,2"; a device may be named twice -- two contexts on one GPU, which is how the
// sharding is tested on a one-GPU box).  CUDA_VISIBLE_DEVICES restricts what is visible as usual.
std::vector<int> device_list()
{
    std::vector<int> devs;
    if (const char *e = getenv("AG2_DEVICES")) {
        devs = ag2host::parse_device_list(e);
    } else {
        int n = 0;
        if (ag2_device_count(&n) == AG2_OK)
            for (int d = 0; d < n; ++d) devs.push_back(d);
    }
    return devs;
}

// load_fastq's batches (mecat2ref_impl_large.cpp:1965-1991: up to SVM reads / MAXSTR characters, plus the record that ended
// the loop) straight from the conversion of the reads file: chang_fastqfile's pass over the FASTA / FASTQ input runs on
// its own thread, writes <wrk>/0.fq as ever, and hands every batch to the mapping as soon as it is complete -- the reference
// (and round 1) finish the conversion first and then parse 0.fq again.
struct ReadBatch {
    std::string bases;
    std::vector<int64_t> offs;
    std::vector<int> ids;
};

class ReadFeeder {
public:
    ReadFeeder(const char *in_path, const std::string &out_path)
    {
        th_ = std::thread([this, in_path, out_path] {
            auto cur = std::make_unique<ReadBatch>();
            cur->offs.assign(1, 0);
            long sum = 0;
            auto push = [&](std::unique_ptr<ReadBatch> b) {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return q_.size() < 2; });
                q_.push_back(std::move(b));
                cv_.notify_all();
            };
            const int n = convert_to_fq(in_path, out_path, [&](int id, const char *seq, size_t len) {
                const bool within = (int)cur->ids.size() < kSVM && sum < kMAXSTR;
                cur->bases.append(seq, len);
                cur->offs.push_back((int64_t)cur->bases.size());
                cur->ids.push_back(id);
                sum += (long)len + 1;
                if (!within) {
                    push(std::move(cur));
                    cur = std::make_unique<ReadBatch>();
                    cur->offs.assign(1, 0);
                    sum = 0;
                }
            });
            if (!cur->ids.empty()) push(std::move(cur));
            std::lock_guard<std::mutex> g(m_);
            count_ = n;
            done_ = true;
            cv_.notify_all();
        });
    }
    // next batch, or null at the end of the file
    std::unique_ptr<ReadBatch> next()
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return !q_.empty() || done_; });
        if (q_.empty()) return nullptr;
        std::unique_ptr<ReadBatch> b = std::move(q_.front());
        q_.erase(q_.begin());
        cv_.notify_all();
        return b;
    }
    // number of reads; waits for the conversion to end (batches nobody took are dropped)
    int finish()
    {
        {
            std::unique_lock<std::mutex> g(m_);
            while (!done_) {
                q_.clear();
                cv_.notify_all();
                cv_.wait_for(g, std::chrono::milliseconds(20));
            }
        }
        if (th_.joinable()) th_.join();
        return count_;
    }
    ~ReadFeeder() { finish(); }

private:
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_;
    std::vector<std::unique_ptr<ReadBatch>> q_;
    bool done_ = false;
    int count_ = 0;
};

// Page-locked result buffers (ag2_host_alloc), recycled between batches: a fresh std::vector of a gigabyte is a second of
// page faults and zero-filling, and device copies into pageable memory run at a fraction of the PCIe rate.
class PinnedPool {
public:
    struct Block {
        char *p = nullptr;
        size_t cap = 0;
    };
    Block take(size_t bytes)
    {
        {
            std::lock_guard<std::mutex> g(m_);
            size_t best = free_.size();
            for (size_t i = 0; i < free_.size(); ++i)
                if (free_[i].cap >= bytes && (best == free_.size() || free_[i].cap < free_[best].cap)) best = i;
            if (best != free_.size()) {
                Block b = free_[best];
                free_.erase(free_.begin() + (long)best);
                return b;
            }
        }
        Block b;
        b.cap = bytes + bytes / 8 + 4096;
        void *q = nullptr;
        if (ag2_host_alloc(b.cap, &q) != AG2_OK) {
            fprintf(stderr, "mecat2ref (aligngraph2_b200): cannot allocate %zu bytes of host memory\n", b.cap);
            fflush(stderr);
            _exit(1);
        }
        b.p = (char *)q;
        return b;
    }
    void give(Block b)
    {
        if (!b.p) return;
        std::lock_guard<std::mutex> g(m_);
        free_.push_back(b);
    }
    ~PinnedPool()
    {
        for (Block &b : free_) ag2_host_free(b.p);
    }

private:
    std::mutex m_;
    std::vector<Block> free_;
};

struct DeviceShard {   // one GPU's context and what it produced for the current batch
    ag2_ctx *ctx = nullptr;
    int64_t lo = 0, hi = 0;   // its contiguous range of the batch's reads
    std::vector<int64_t> offs;
    PinnedPool::Block rec, qaln, saln;   // ag2_record[n_rec], the two string pools
    int64_t n_rec = 0;
    int rc = AG2_OK;
    const char *what = "";
};

template <class F> void on_every_device(std::vector<DeviceShard> &sh, F f)
{
    if (sh.size() == 1) {
        f(sh[0]);
    } else {
        std::vector<std::thread> th;
        for (DeviceShard &d : sh) th.emplace_back([&d, &f] { f(d); });
        for (std::thread &t : th) t.join();
    }
    for (DeviceShard &d : sh)
        if (d.rc != AG2_OK) die_ag2(d.ctx, d.what, d.rc);
}

// the mapping part of meap_ref_impl_large (:1994-2149) on the GPU; returns seconds {read index, ref index, mapping}
void map_on_gpu(const Options &o, const std::string &ref_seq, double secs[3], ResultWriter &rw, ReadFeeder &feeder)
{
    const std::vector<int> devs = device_list();
    std::vector<DeviceShard> sh;
    for (size_t k = 0; k < devs.size(); ++k) {   // the devices that initialise are used; the run fails only if none does
        ag2_ctx *ctx = nullptr;
        const int rc = ag2_ctx_create(devs[k], &ctx);
        if (rc != AG2_OK) {
            fprintf(stderr, "mecat2ref (aligngraph2_b200): device %d is not usable (%d), going on without it\n", devs[k], rc);
            continue;
        }
        sh.emplace_back();
        sh.back().ctx = ctx;
    }
    if (sh.empty()) {
        fprintf(stderr, "mecat2ref (aligngraph2_b200): no usable CUDA device; there is no CPU path\n");
        fflush(stderr);
        _exit(1);   // the conversion thread is still writing 0.fq: no stdio teardown under it
    }
    const size_t ndev = sh.size();
    double t0 = now_sec();
    on_every_device(sh, [&](DeviceShard &d) {
        d.what = "ag2_ref_load";
        d.rc = ag2_ref_load(d.ctx, ref_seq.data(), (int64_t)ref_seq.size());
    });
    secs[1] = now_sec() - t0;
    const std::string wrk = o.wrk_dir;
    for (int t = 2; t <= o.num_cores; ++t) fclose(fopen((wrk + "/" + std::to_string(t) + ".r").c_str(), "w"));
    for (int t = 1; t <= o.num_cores; ++t) fclose(fopen((wrk + "/ref" + std::to_string(t) + ".r").c_str(), "w"));

    // the batches come from the conversion thread (ReadFeeder); while the GPUs map one and the writer prints the one before,
    // the feeder is already parsing the next
    PinnedPool pool;
    std::thread writer;
    bool first_batch = true;
    secs[0] = secs[2] = 0;
    for (;;) {
        std::unique_ptr<ReadBatch> cur_p = feeder.next();
        if (!cur_p) break;
        const ReadBatch &cur = *cur_p;
        const std::string &bases = cur.bases;
        const std::vector<int64_t> &offs = cur.offs;
        const std::vector<int> &ids = cur.ids;
        const int64_t n_reads = (int64_t)ids.size();
        t0 = now_sec();
        if (first_batch) {
            // build_read_index uses the first <= 100 000 reads of the file; the index is built once (:2017-2033), on every
            // GPU from the same reads (the whole first batch), so every GPU votes and seeds with the same tables
            on_every_device(sh, [&](DeviceShard &d) {
                d.what = "ag2_reads_load";
                if ((d.rc = ag2_reads_load(d.ctx, bases.data(), offs.data(), n_reads)) != AG2_OK) return;
                d.what = "ag2_index_build";
                d.rc = ag2_index_build(d.ctx, o.block, o.alpha, o.beta);
            });
            secs[0] = now_sec() - t0;
            t0 = now_sec();
        }
        // contiguous read ranges, balanced by bases: concatenating the GPUs' records in device order keeps the file order
        const std::vector<std::pair<int64_t, int64_t>> ranges = ag2host::split_by_bases(offs, ndev);
        for (size_t k = 0; k < ndev; ++k) {
            sh[k].lo = ranges[k].first;
            sh[k].hi = ranges[k].second;
        }
        on_every_device(sh, [&](DeviceShard &d) {
            d.n_rec = 0;
            if (d.hi <= d.lo) return;
            if (!(first_batch && ndev == 1)) {   // (one GPU, first batch: its reads are the batch that is loaded already)
                d.offs.resize((size_t)(d.hi - d.lo) + 1);
                for (int64_t r = d.lo; r <= d.hi; ++r) d.offs[(size_t)(r - d.lo)] = offs[(size_t)r] - offs[(size_t)d.lo];
                d.what = "ag2_reads_load";
                if ((d.rc = ag2_reads_load(d.ctx, bases.data() + offs[(size_t)d.lo], d.offs.data(), d.hi - d.lo)) != AG2_OK) return;
            }
            int64_t used = 0;
            d.what = "ag2_map_reads";
            if ((d.rc = ag2_map_reads(d.ctx, o.num_candidates, o.num_output, &d.n_rec)) != AG2_OK) return;
            d.rec = pool.take(((size_t)d.n_rec + 1) * sizeof(ag2_record));
            d.what = "ag2_map_fetch";
            if ((d.rc = ag2_map_fetch(d.ctx, (ag2_record *)d.rec.p, nullptr, nullptr, 0, &used)) != AG2_OK) return;
            d.qaln = pool.take((size_t)used + 1);
            d.saln = pool.take((size_t)used + 1);
            d.rc = ag2_map_fetch(d.ctx, (ag2_record *)d.rec.p, d.qaln.p, d.saln.p, used, &used);
        });
        first_batch = false;
        secs[2] += now_sec() - t0;
        // the batch's records go to the writer (1.r, -o, -p from memory) on its own thread while the next batch is mapped
        struct BatchResult {
            std::vector<PinnedPool::Block> rec, q, s;
            std::vector<int64_t> lo, n_rec;
            std::vector<int> ids;
        };
        auto res = std::make_shared<BatchResult>();
        for (DeviceShard &d : sh) {
            res->rec.push_back(d.rec);
            res->q.push_back(d.qaln);
            res->s.push_back(d.saln);
            res->lo.push_back(d.lo);
            res->n_rec.push_back(d.n_rec);
            d.rec = d.qaln = d.saln = PinnedPool::Block();
        }
        res->ids = ids;
        if (writer.joinable()) writer.join();
        writer = std::thread([res, &rw, &pool] {
            std::vector<RecView> views;
            size_t total = 0;
            for (int64_t n : res->n_rec) total += (size_t)n;
            views.reserve(total);
            for (size_t k = 0; k < res->rec.size(); ++k)
                for (int64_t i = 0; i < res->n_rec[k]; ++i) { // output_temp_result (output.cpp:237-251)
                    const ag2_record &r = ((const ag2_record *)res->rec[k].p)[(size_t)i];
                    RecView v;
                    v.read_id = res->ids[(size_t)(res->lo[k] + r.read)];
                    v.read_dir = r.strand ? 'R' : 'F';
                    v.vscore = r.vscore, v.qb = r.qb, v.qe = r.qe, v.qs = r.qs;
                    v.sb = (long)r.sb, v.se = (long)r.se;
                    v.qmap = res->q[k].p + r.aln_off;
                    v.smap = res->s[k].p + r.aln_off;
                    v.len = r.aln_len;
                    views.push_back(v);
                }
            rw.add_batch(views);
            for (size_t k = 0; k < res->rec.size(); ++k) {   // the buffers go back for the batch after next
                pool.give(res->rec[k]);
                pool.give(res->q[k]);
                pool.give(res->s[k]);
            }
        });
    }
    if (writer.joinable()) writer.join();
    for (DeviceShard &d : sh) ag2_ctx_destroy(d.ctx);
}

} // namespace

int main(int argc, char **argv)
{
    prog_name = argv[0];
    const double t_start = now_sec();
    Options o;
    if (parse_options(argc, argv, o) == -1) {
        print_usage();
        return 1;
    }
    const std::string wrk = o.wrk_dir;
    const bool trace = getenv("AG2_TRACE") != nullptr;   // host timeline on stderr
    double t_stage = now_sec();
    auto stage_done = [&](const char *what) {
        if (trace) fprintf(stderr, "[mecat2ref host] %-28s %8.3f s\n", what, now_sec() - t_stage);
        t_stage = now_sec();
    };
    // chang_fastqfile on its own thread: <wrk>/0.fq is written as ever while the batches already go to the GPUs
    const bool skip_map = getenv("AG2_SKIP_MAP") != nullptr;
    std::unique_ptr<ReadFeeder> feeder;
    int readcount = 0;
    if (skip_map) readcount = convert_to_fq(o.reads, wrk + "/0.fq");
    else feeder = std::make_unique<ReadFeeder>(o.reads, wrk + "/0.fq");
    const int refcount = convert_to_fq(o.reference, wrk + "/ref.fq");
    stage_done("reference -> ref.fq (reads -> 0.fq on its own thread)");
    auto write_config = [&]() {
        FILE *cfg = fopen("config.txt", "w");
        if (!cfg) { fprintf(stderr, "failed to open config.txt for writing\n"); exit(1); }
        fprintf(cfg, "%s\n%s\n%s\n%s\n%s\n%d\t%d\n%d\n", o.wrk_dir, o.reference, o.reads, o.output, o.refoutput, o.num_cores, readcount, refcount);
        fclose(cfg);
    };
    if (skip_map) write_config();
    printf("first task is sucess\n");
    double secs[3] = {0, 0, 0};
    bool wrote_results = false;
    {
        std::string ref_seq;
        const double t0 = now_sec();
        load_reference(o.reference, wrk, ref_seq);
        const double t_load = now_sec() - t0;
        stage_done("reference, chrindex.txt");
        if (!getenv("AG2_SKIP_MAP")) {
            // 1.r, -o and -p are written from the records in memory while the mapping goes on
            const std::vector<ChrInfo> chr0 = read_chrindex(wrk);
            ResultWriter rw(o, chr0, argc, argv);
            map_on_gpu(o, ref_seq, secs, rw, *feeder);
            rw.finish();
            wrote_results = true;
            readcount = feeder->finish();
            write_config();   // (the reference writes it before the mapping; nothing reads it in between)
        }
        secs[1] += t_load;
        stage_done("mapping (read, map, 1.r, -o, -p)");
    }
    {
        FILE *cfg = fopen("config.txt", "a");
        fprintf(cfg, "The Building read Index Time: %f sec\n", secs[0]);
        fprintf(cfg, "The Building  Reference  Index Time: %f sec\n", secs[1]);
        fprintf(cfg, "The Mapping Time: %f sec\n", secs[2]);
        fclose(cfg);
    }
    const std::vector<ChrInfo> chr = read_chrindex(wrk);
    // Without a mapping run (AG2_SKIP_MAP: the thread files are given) -o / -p come from the thread files as in the reference.
    // The two passes are independent (own output file, no shared state): side by side, unless -o and -p name the same file,
    // which the reference would write one after the other
    if (wrote_results) {
    } else if (strcmp(o.output, o.refoutput) != 0) {
        std::thread combine([&] { result_combine(o, chr, argc, argv); });
        polish_result(o, chr, argc, argv);
        combine.join();
        stage_done("result_combine + polish_result");
    } else {
        result_combine(o, chr, argc, argv);
        stage_done("result_combine (-o)");
        polish_result(o, chr, argc, argv);
        stage_done("polish_result (-p)");
    }
    {
        FILE *cfg = fopen("config.txt", "a");
        fprintf(cfg, "The total Time : %f sec\n", now_sec() - t_start);
        fclose(cfg);
    }
    const std::string cmd = std::string("cp -r config.txt \"") + o.refoutput + ".config\"";
    const int st = system(cmd.c_str());
    if (st != 0) {
        fprintf(stderr, "[main, %u] system() error. Error code is %d.\n", __LINE__, st);
        return 1;
    }
    return EXIT_SUCCESS;
}
