// Prints what aligngraph2_b200/host/shard_split.h decides, for tests/test_host_shard.py (test infrastructure).
//   probe devices "<list>"            -> the parsed device ids, space separated
//   probe split <parts> <len>...      -> "lo hi" per part for reads of the given lengths
#include "../../aligngraph2_b200/host/shard_split.h"

#include <cstdio>
#include <cstring>

int main(int argc, char **argv)
{
    if (argc >= 3 && !strcmp(argv[1], "devices")) {
        for (int d : ag2host::parse_device_list(argv[2])) printf("%d ", d);
        printf("\n");
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "split")) {
        const size_t parts = (size_t)atol(argv[2]);
        std::vector<int64_t> offs(1, 0);
        for (int i = 3; i < argc; ++i) offs.push_back(offs.back() + atoll(argv[i]));
        for (const auto &r : ag2host::split_by_bases(offs, parts)) printf("%lld %lld\n", (long long)r.first, (long long)r.second);
        return 0;
    }
    return 2;
}
