// gpu_commands.cpp -- TEST INFRASTRUCTURE (oracle/): the reference-side binding of libplassgpu.so, compiled.
//
// This translation unit is what INTEGRATION.md section 3 describes: the reference's own tool table (src/plass.cpp, included
// from where it lies under the reference tree -- nothing is copied) plus three Command rows, registered in front of the
// existing ones so that Application.cpp:24-36 resolves `kmermatcher`, `rescorediagonal` and `assembleresults` to the GPU
// shims.  The shims use the REFERENCE's classes for everything around the hot path -- Parameters for the flags, DBReader for
// the input DBs, DBWriter + QueryMatcher::prefilterHitToBuffer / Matcher::resultToBuffer for the outputs -- and call the
// C ABI of include/plassgpu.h for the arithmetic.  Linked against oracle/_ref/libmmseqs-framework.a by oracle/ref_build.mk
// into oracle/_ref/bin/plass_gpu_shim; tests/test_gpu_dropin.py runs the reference's assemble.sh through it.
#include PLASS_TOOL_CPP          // the reference's src/plass.cpp: tool globals + std::vector<Command> commands

#include "DBReader.h"
#include "DBWriter.h"
#include "Debug.h"
#include "Matcher.h"
#include "QueryMatcher.h"
#include "Util.h"
#include "kmermatcher.h"

#include "plassgpu.h"

#include <climits>
#include <string>
#include <vector>

namespace {

pg_context *gpuContext() {
    static pg_context *ctx = NULL;
    if (ctx == NULL && pg_init(0, &ctx) != 0) {
        Debug(Debug::ERROR) << pg_last_error() << "\n";
        EXIT(EXIT_FAILURE);
    }
    return ctx;
}

void check(int rc) {
    if (rc != 0) {
        Debug(Debug::ERROR) << pg_last_error() << "\n";
        EXIT(EXIT_FAILURE);
    }
}

// DBReader view -> device DB.  Entries are packed in index (= key) order; DBReader::getData hides split data files.
struct HostSeqDb {
    std::string data;
    std::vector<uint64_t> offsets;
    std::vector<uint32_t> lens, keys;
    int dbtype;
};

pg_seqdb *uploadSequences(DBReader<unsigned int> &reader, HostSeqDb &h) {
    const size_t n = reader.getSize();
    h.offsets.resize(n); h.lens.resize(n); h.keys.resize(n);
    size_t total = 0;
    for (size_t i = 0; i < n; i++) {
        h.offsets[i] = total;
        h.lens[i] = (uint32_t) reader.getEntryLen(i);
        h.keys[i] = reader.getDbKey(i);
        total += h.lens[i];
    }
    h.data.resize(total);
    for (size_t i = 0; i < n; i++) {
        memcpy(&h.data[h.offsets[i]], reader.getData(i, 0), h.lens[i]);
    }
    h.dbtype = Parameters::isEqualDbtype(reader.getDbtype(), Parameters::DBTYPE_NUCLEOTIDES) ? PG_DBTYPE_NUCLEOTIDES : PG_DBTYPE_AMINO_ACIDS;
    pg_seqdb_view v;
    v.data = h.data.data(); v.data_bytes = h.data.size();
    v.offsets = h.offsets.data(); v.lens = h.lens.data(); v.keys = h.keys.data(); v.n = n; v.dbtype = h.dbtype;
    pg_seqdb *db = NULL;
    check(pg_seqdb_upload(gpuContext(), &v, &db));
    return db;
}

int gpu_kmermatcher(int argc, const char **argv, const Command &command) {
    Parameters &par = Parameters::getInstance();
    setLinearFilterDefault(&par);
    par.parseParameters(argc, argv, command, true, 0, MMseqsParameter::COMMAND_CLUSTLINEAR);
    DBReader<unsigned int> seqDbr(par.db1.c_str(), par.db1Index.c_str(), par.threads, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    seqDbr.open(DBReader<unsigned int>::NOSORT);
    const bool nucl = Parameters::isEqualDbtype(seqDbr.getDbtype(), Parameters::DBTYPE_NUCLEOTIDES);
    if (par.maskMode != 0 || par.maskLowerCaseMode != 0 || par.spacedKmer != 0 || par.adjustKmerLength != 0 || par.compressed != 0) {
        Debug(Debug::ERROR) << "GPU kmermatcher: masking, spaced k-mers, k-mer length adjustment and compressed DBs are not supported\n";
        EXIT(EXIT_FAILURE);
    }
    HostSeqDb host;
    pg_seqdb *db = uploadSequences(seqDbr, host);
    pg_km_params p;
    p.kmer_size = (int) par.kmerSize;
    p.alph_size = nucl ? par.alphabetSize.nucleotides : par.alphabetSize.aminoacids;
    p.kmers_per_seq = (int) par.kmersPerSequence;
    p.kmers_per_seq_scale = nucl ? par.kmersPerSequenceScale.nucleotides : par.kmersPerSequenceScale.aminoacids;
    p.hash_shift = (int) par.hashShift;
    p.include_only_extendable = par.includeOnlyExtendable ? 1 : 0;
    p.ignore_multi_kmer = par.ignoreMultiKmer ? 1 : 0;
    p.cov_mode = par.covMode;
    p.cov_thr = par.covThr;
    p.hash_start = 0; p.hash_end = 65535;
    check(pg_set_split_memory_limit(gpuContext(), par.splitMemoryLimit));
    pg_hit *hits = NULL; uint64_t nHits = 0;
    check(pg_kmermatch(gpuContext(), db, &p, &hits, &nHits));
    DBWriter dbw(par.db2.c_str(), par.db2Index.c_str(), 1, par.compressed, nucl ? Parameters::DBTYPE_PREFILTER_REV_RES : Parameters::DBTYPE_PREFILTER_RES);
    dbw.open();
    std::string block;
    char buffer[100];
    uint64_t h = 0;
    for (size_t i = 0; i < host.keys.size(); i++) {
        const unsigned int key = host.keys[i];
        block.clear();
        hit_t self; self.seqId = key; self.prefScore = 0; self.diagonal = 0;
        block.append(buffer, QueryMatcher::prefilterHitToBuffer(buffer, self));
        while (h < nHits && hits[h].rep < key) h++;
        for (; h < nHits && hits[h].rep == key; h++) {
            hit_t hit; hit.seqId = hits[h].target; hit.prefScore = hits[h].score; hit.diagonal = (unsigned short) hits[h].diag;
            block.append(buffer, QueryMatcher::prefilterHitToBuffer(buffer, hit));
        }
        dbw.writeData(block.c_str(), block.size(), key, 0);
    }
    dbw.close(false, false);
    seqDbr.close();
    pg_free_host(hits);
    pg_seqdb_free(gpuContext(), db);
    return EXIT_SUCCESS;
}

int gpu_rescorediagonal(int argc, const char **argv, const Command &command) {
    Parameters &par = Parameters::getInstance();
    par.parseParameters(argc, argv, command, true, 0, 0);
    if (par.rescoreMode != Parameters::RESCORE_MODE_END_TO_END_ALIGNMENT || par.db1 != par.db2 || par.compressed != 0 || par.sortResults != 0 || par.filterHits) {
        Debug(Debug::ERROR) << "GPU rescorediagonal: only --rescore-mode 3 on query DB == target DB, unsorted, unfiltered, uncompressed\n";
        EXIT(EXIT_FAILURE);
    }
    DBReader<unsigned int> seqDbr(par.db1.c_str(), par.db1Index.c_str(), par.threads, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    seqDbr.open(DBReader<unsigned int>::NOSORT);
    DBReader<unsigned int> prefDbr(par.db3.c_str(), par.db3Index.c_str(), par.threads, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    prefDbr.open(DBReader<unsigned int>::NOSORT);
    HostSeqDb host;
    pg_seqdb *db = uploadSequences(seqDbr, host);
    // prefilter lines in query-key order, each block without its self line
    std::vector<pg_hit> hits;
    for (size_t i = 0; i < host.keys.size(); i++) {
        const size_t id = prefDbr.getId(host.keys[i]);
        if (id == UINT_MAX) continue;
        char *data = prefDbr.getData(id, 0);
        bool first = true;
        while (*data != '\0') {
            hit_t hit = QueryMatcher::parsePrefilterHit(data);
            if (!first) {
                pg_hit h; h.rep = host.keys[i]; h.target = hit.seqId; h.score = hit.prefScore; h.diag = (int32_t) (short) hit.diagonal;
                hits.push_back(h);
            }
            first = false;
            data = Util::skipLine(data);
        }
    }
    pg_rs_params p;
    p.rescore_mode = 3; p.seq_id_thr = par.seqIdThr; p.eval_thr = par.evalThr; p.cov_mode = par.covMode; p.cov_thr = par.covThr;
    p.aln_len_thr = par.alnLenThr; p.seq_id_mode = par.seqIdMode;
    pg_aln *alns = NULL; uint64_t nAlns = 0;
    check(pg_rescore(gpuContext(), db, hits.data(), hits.size(), &p, &alns, &nAlns));
    DBWriter dbw(par.db4.c_str(), par.db4Index.c_str(), 1, par.compressed, Parameters::DBTYPE_ALIGNMENT_RES);
    dbw.open();
    std::string block;
    char buffer[1024];
    uint64_t a = 0;
    for (size_t i = 0; i < host.keys.size(); i++) {
        const unsigned int key = host.keys[i];
        block.clear();
        while (a < nAlns && alns[a].query < key) a++;
        for (; a < nAlns && alns[a].query == key; a++) {
            const pg_aln &r = alns[a];
            const unsigned int alnLen = (unsigned int) (std::max(abs(r.q_end - r.q_start), abs(r.db_end - r.db_start)) + 1);
            Matcher::result_t res(r.target, r.bits, 0.0f, 0.0f, r.seq_id, r.evalue, alnLen, r.q_start, r.q_end, (unsigned int) r.q_len,
                                  r.db_start, r.db_end, (unsigned int) r.db_len, std::string());
            block.append(buffer, Matcher::resultToBuffer(buffer, res, false));
        }
        dbw.writeData(block.c_str(), block.size(), key, 0);
    }
    dbw.close();
    seqDbr.close(); prefDbr.close();
    pg_free_host(alns);
    pg_seqdb_free(gpuContext(), db);
    return EXIT_SUCCESS;
}

int gpu_assembleresults(int argc, const char **argv, const Command &command) {
    LocalParameters &par = LocalParameters::getLocalInstance();
    par.parseParameters(argc, argv, command, true, 0, 0);
    DBReader<unsigned int> seqDbr(par.db1.c_str(), par.db1Index.c_str(), par.threads, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    seqDbr.open(DBReader<unsigned int>::NOSORT);
    DBReader<unsigned int> alnDbr(par.db2.c_str(), par.db2Index.c_str(), par.threads, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    alnDbr.open(DBReader<unsigned int>::NOSORT);
    HostSeqDb host;
    pg_seqdb *db = uploadSequences(seqDbr, host);
    std::vector<pg_aln> alns;
    std::vector<Matcher::result_t> parsed;
    for (size_t i = 0; i < host.keys.size(); i++) {
        const size_t id = alnDbr.getId(host.keys[i]);
        if (id == UINT_MAX) continue;
        parsed.clear();
        Matcher::readAlignmentResults(parsed, alnDbr.getData(id, 0), false);
        for (size_t j = 0; j < parsed.size(); j++) {
            const Matcher::result_t &r = parsed[j];
            pg_aln a;
            a.query = host.keys[i]; a.target = r.dbKey; a.bits = r.score; a.seq_id = r.seqId; a.evalue = r.eval;
            a.q_start = r.qStartPos; a.q_end = r.qEndPos; a.q_len = (int32_t) r.qLen;
            a.db_start = r.dbStartPos; a.db_end = r.dbEndPos; a.db_len = (int32_t) r.dbLen;
            alns.push_back(a);
        }
    }
    pg_ex_params p;
    p.seq_id_thr = par.seqIdThr; p.max_seq_len = (int) par.maxSeqLen; p.keep_target = par.keepTarget ? 1 : 0; p.rescore_mode = par.rescoreMode;
    pg_seqdb *out = NULL;
    check(pg_extend(gpuContext(), db, alns.data(), alns.size(), &p, &out, NULL));
    char *data; uint64_t bytes, *offs, n; uint32_t *lens, *keys;
    check(pg_seqdb_download(gpuContext(), out, &data, &bytes, &offs, &lens, &keys, &n));
    DBWriter dbw(par.db3.c_str(), par.db3Index.c_str(), 1, par.compressed, seqDbr.getDbtype());
    dbw.open();
    for (uint64_t i = 0; i < n; i++) {
        dbw.writeData(data + offs[i], lens[i] - 1, keys[i], 0);
    }
    dbw.close(true);
    seqDbr.close(); alnDbr.close();
    pg_free_host(data); pg_free_host(offs); pg_free_host(lens); pg_free_host(keys);
    pg_seqdb_free(gpuContext(), out);
    pg_seqdb_free(gpuContext(), db);
    return EXIT_SUCCESS;
}

// the three rows go IN FRONT of the tool's table: its own `assembleresults` row and the framework's `kmermatcher` /
// `rescorediagonal` rows (baseCommands, searched second) are shadowed without touching a script
struct RegisterGpuCommands {
    RegisterGpuCommands() {
        Parameters &par = Parameters::getInstance();
        std::vector<Command> rows = {
            {"kmermatcher", gpu_kmermatcher, &par.kmermatcher, COMMAND_HIDDEN, "B200 k-mer matcher (libplassgpu.so)", NULL, "", "<i:sequenceDB> <o:prefilterDB>",
             CITATION_PLASS, {{"sequenceDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::sequenceDb},
                              {"prefilterDB", DbType::ACCESS_MODE_OUTPUT, DbType::NEED_DATA, &DbValidator::prefilterDb}}},
            {"rescorediagonal", gpu_rescorediagonal, &par.rescorediagonal, COMMAND_HIDDEN, "B200 ungapped diagonal rescoring (libplassgpu.so)", NULL, "",
             "<i:queryDB> <i:targetDB> <i:prefilterDB> <o:resultDB>",
             CITATION_PLASS, {{"queryDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::sequenceDb},
                              {"targetDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::sequenceDb},
                              {"resultDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::resultDb},
                              {"alignmentDB", DbType::ACCESS_MODE_OUTPUT, DbType::NEED_DATA, &DbValidator::alignmentDb}}},
            {"assembleresults", gpu_assembleresults, &localPar.assembleresults, COMMAND_HIDDEN, "B200 greedy extension (libplassgpu.so)", NULL, "",
             "<i:sequenceDB> <i:alnResult> <o:reprSeqDB>",
             CITATION_PLASS, {{"sequenceDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::sequenceDb},
                              {"alnResult", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::alignmentDb},
                              {"reprSeqDB", DbType::ACCESS_MODE_OUTPUT, DbType::NEED_DATA, &DbValidator::sequenceDb}}},
        };
        // (Command has const members: no assignment, so the table is rebuilt by copy construction and swapped in)
        std::vector<Command> merged;
        merged.reserve(rows.size() + commands.size());
        for (size_t i = 0; i < rows.size(); i++) merged.push_back(rows[i]);
        for (size_t i = 0; i < commands.size(); i++) merged.push_back(commands[i]);
        commands.swap(merged);
    }
} registerGpuCommands;

}  // namespace
