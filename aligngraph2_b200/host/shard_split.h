// shard_split.h -- how the drop-in executables spread a batch of reads over GPUs (contiguous read ranges,
// no data-path collective).  Pure host logic, no CUDA: unit-tested on the CPU (tests/test_host_shard.py).
#ifndef AG2_HOST_SHARD_SPLIT_H
#define AG2_HOST_SHARD_SPLIT_H

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <utility>
#include <vector>

namespace ag2host {

// "0,1,2" -> {0, 1, 2}.  A device may be named more than once (several contexts on one GPU: how the sharding is tested on
// a one-GPU box).  Parsing stops at the first thing that is not a number or a comma.
inline std::vector<int> parse_device_list(const char *e)
{
    std::vector<int> devs;
    for (const char *p = e; p && *p;) {
        char *q = nullptr;
        const long d = strtol(p, &q, 10);
        if (q == p) break;
        devs.push_back((int)d);
        p = (*q == ',') ? q + 1 : q;
    }
    return devs;
}

// Cuts reads [0, n) with base offsets offs[0..n] into `parts` contiguous ranges [lo, hi) balanced by BASES (reads differ in
// length); ranges may be empty when there are fewer reads than parts.  Concatenating the parts in order gives back the
// batch in read order, which is what keeps the output files independent of the number of GPUs.
inline std::vector<std::pair<int64_t, int64_t>> split_by_bases(const std::vector<int64_t> &offs, size_t parts)
{
    const int64_t n = (int64_t)offs.size() - 1;
    std::vector<std::pair<int64_t, int64_t>> out(parts);
    for (size_t k = 0; k < parts; ++k) {
        const int64_t want = (int64_t)((double)offs[(size_t)n] * (double)(k + 1) / (double)parts);
        const int64_t lo = k ? out[k - 1].second : 0;
        int64_t hi = k + 1 == parts ? n : std::max<int64_t>(lo, std::upper_bound(offs.begin(), offs.end(), want) - offs.begin() - 1);
        hi = std::min(hi, n);
        out[k] = {lo, hi};
    }
    return out;
}

} // namespace ag2host

#endif
