/*
 * ref_pagraph_dump.cpp -- drives the UNMODIFIED PAGraph sources up to the end of the A-Bruijn build and dumps
 * the graph (TEST INFRASTRUCTURE ONLY).
 *
 * Compiled by oracle/Makefile together with every src/tools/ translation unit of the reference (read in place from
 * /root/reference) into oracle/_ref/pagraph_dump.  It adds no algorithm: main() makes the calls that
 * PAGraph/src/main/pagraph.cpp:69-243 (run2) makes, with the same constants, in the same order, up to
 * PositionProcessor::process(); where run2 goes on to PAssembly::testTravel5 this program instead writes every
 * vertex of the dense table.  The unmodified `pagraph` binary (oracle/_ref/pagraph) cannot be used to pin rows
 * B2-B8 because it only ever prints the vertices that end up on a path.
 *
 *   pagraph_dump <thread> <kmer.bin> <ctg.fasta> <ref.fasta> <pre dir> <ctg-to-ref aln> <epsilon> <cov> <out.txt>
 *                [<travel threads> <travel out dir> [<minLen>]]
 *
 * With the optional arguments the program goes on exactly as run2 does (pagraph.cpp:225-269): PAssembly::testTravel5
 * on every block and contig.txt -- but with its own thread count, so that the traversal's `min(threadNum, 8)` start
 * vertices (PAlgorithm.cpp:146) can be pinned on the deterministic `-t 1` graph (row B9; the unmodified `pagraph -t 8`
 * would also build the graph with 8 threads, which is not reproducible).
 *
 * Dump format (one line per vertex that holds anything; vertices in dense-index order):
 *   #config <n> <ref name>
 *   V <dense idx> <k-mer code> P <npos> {<ctg>,<ref>,<count>}... E <nedge> {<to dense idx>,<step>}...
 *
 * The one liberty taken: `#define private public` around the reference headers, because PABruijnGraph keeps its
 * dense table private and exposes no edge accessor.  Access specifiers do not change the class layout.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <map>
#include <memory>
#include <mutex>
#include <regex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#define private public
#include "align/MummerAlignDatabaseV2.hpp"
#include "align/MecatAlignDatabase.hpp"
#include "seq/AutoSeqDatabase.hpp"
#include "graph/PABruijnGraph.hpp"
#include "align/Aligner.hpp"
#include "position/PositionProcessor.hpp"
#include "kmer/FileKmerIterator.hpp"
#include "graph/PAssembly.hpp"
#include "position/PositionMapper.hpp"
#undef private

struct Cfg {
    std::string ref, reads, ctgAln, refAln;
    std::vector<std::pair<std::string, bool>> contigs;
};

/* same file grammar as loadFromConfig, PAGraph/src/main/pagraph.cpp:29-49 */
static std::vector<Cfg> load_cfg(const std::string& path)
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

int main(int argc, char** argv)
{
    if (argc != 10 && argc != 12 && argc != 13) {
        fprintf(stderr, "usage: pagraph_dump t kmer ctg ref pre aln eps cov out\n");
        return 2;
    }
    unsigned threadNum = (unsigned)atoi(argv[1]);
    std::string kmerPath = argv[2], contigsPath = argv[3], referencesPath = argv[4], inputDir = argv[5],
                contigToRefPath = argv[6];
    std::size_t posError = (std::size_t)atoll(argv[7]), covFilter = (std::size_t)atoll(argv[8]);
    FILE* out = fopen(argv[9], "w");
    if (!out) return 2;
    const bool travel = argc >= 12;
    unsigned travelThreads = travel ? (unsigned)atoi(argv[10]) : 0;
    std::string outDir = travel ? argv[11] : "";
    std::size_t minLen = argc == 13 ? (std::size_t)atoll(argv[12]) : 50;
    std::unordered_set<std::string> okCtg;

    auto configs = load_cfg(inputDir + "/config.txt");
    auto pKmerIt = std::make_shared<FileKmerIterator>(kmerPath);
    auto pContigDB = std::make_shared<AutoSeqDatabase>(contigsPath);
    auto pRefDB = std::make_shared<AutoSeqDatabase>(referencesPath);
    auto pContigToRef = std::make_shared<MummerAlignDatabaseV2>(contigToRefPath);
    auto pPaGraph = std::make_shared<PABruijnGraph>(*pKmerIt, threadNum);

    int n = 0;
    for (auto& config : configs) {
        pPaGraph->resetAllNodes(threadNum);
        auto pReadDB = std::make_shared<AutoSeqDatabase>(inputDir + "/" + config.reads);
        auto pReadToContig = std::make_shared<MecatAlignDatabase>(inputDir + "/" + config.ctgAln);
        auto pReadToRef = std::make_shared<MecatAlignDatabase>(inputDir + "/" + config.refAln);

        PositionProcessor pp(pPaGraph, pReadDB, pContigDB, pRefDB, pReadToContig, pReadToRef, pContigToRef);
        /* the constants of pagraph.cpp:110-125 */
        pp.setReadToCtgTopK(-1);
        pp.setReadToRefTopK(-1);
        pp.setCtgToRefTopK(-1);
        pp.setOuterSample(3);
        pp.setInnerSample(1);
        pp.setPositionError(posError);
        pp.setReadToCtgRatio(0.35);
        pp.setReadToRefRatio(0.10);
        pp.setCtgToRefRatio(0.00);
        pp.setCtgToRefTotalRatio(0.1);
        pp.setCtgToRefMinLen(50);
        pp.setCovFilter(covFilter);
        pp.setThreadNum(threadNum);
        pp.clearRefFilter(false);
        pp.clearCtgFilter(false);
        pp.setRefFilter(config.ref, true);
        for (auto& ctg : config.contigs) pp.setCtgFilter(ctg.first, ctg.second, true);
        pp.preProcess();
        pp.process();

        fprintf(out, "#config %d %s\n", n++, config.ref.c_str());
        auto& table = *pPaGraph->_pDenseHashTable;
        for (std::size_t v = 0; v < table.size(); ++v) {
            auto& pos = table[v].getAllPositions();
            auto& cnt = table[v].getAllCount();
            auto& ch = table[v].getAllChild();
            if (pos.empty() && ch.empty()) continue;
            fprintf(out, "V %zu %llu P %zu", v, (unsigned long long)pPaGraph->_kmerIndexArr[v], pos.size());
            for (std::size_t i = 0; i < pos.size(); ++i)
                fprintf(out, " %u,%u,%u", pos[i].first, pos[i].second, (unsigned)cnt[i]);
            fprintf(out, " E %zu", ch.size());
            for (auto& e : ch) fprintf(out, " %llu,%d", (unsigned long long)e.first, e.second);
            fputc('\n', out);
        }
        if (travel) {                             /* pagraph.cpp:209-261 */
            std::set<std::pair<std::string, bool>> usedCtg;
            for (auto& ctg : config.contigs) usedCtg.emplace(ctg);
            auto successCtg = PAssembly::testTravel5(outDir, std::to_string(n - 1) + "_", pPaGraph, pReadDB, pContigDB, pRefDB,
                                                     std::make_shared<PositionMapper>(*pContigDB),
                                                     std::make_shared<PositionMapper>(*pRefDB), usedCtg, posError * 2, 0.15, 0.90,
                                                     minLen, travelThreads);
            for (auto& success : successCtg) okCtg.emplace(success.first);
        }
    }
    if (travel) {
        std::ofstream ctgList(outDir + "/contig.txt");
        for (auto& processCtg : okCtg) ctgList << processCtg << std::endl;
    }
    fclose(out);
    return 0;
}
