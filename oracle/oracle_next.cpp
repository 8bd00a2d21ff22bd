/* oracle_next.cpp -- CPU restatement of the components SURVEY.md section 8(f) ranks "next" to the hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Scalar, single-threaded, statement-for-statement restatements of
 *
 *   or_findstart   findassemblystart        src/assembler/findassemblystart.cpp:35-176
 *   or_cyclecheck  cyclecheck               src/assembler/cyclecheck.cpp:71-274
 *   or_extractorfs extractorfs (+ translatenucs --add-orf-stop 1)   mm/util/extractorfs.cpp:20-159, mm/commons/Orf.cpp:127-330,
 *                                            mm/commons/TranslateNucl.h:333-503, mm/util/translatenucs.cpp:14-128
 *
 * Parity status: PINNED -- tests/test_oracle_vs_reference.py checks both against DBs written by the unmodified
 * reference binary (tests/golden/{example_aa,synth_aa}: aln_0 -> corrected_seqs; tests/golden/{synth_nt,long_nt}:
 * assembly_N -> assembly_N_noneCycle / assembly_N_cycle; tests/golden/orf_aa: nucl_reads -> nucl_6f_{start,long}[_h] ->
 * aa_6f_{start,long}).
 */
#include "oracle.h"
#include "oracle_tables.h"

#include <algorithm>
#include <cctype>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

// DBReader::getId (binary search in the key-sorted index); UINT64_MAX if absent
uint64_t idOfKey(const or_seqdb *db, uint32_t key) {
    uint64_t lo = 0, hi = db->n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) / 2;
        if (db->keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < db->n && db->keys[lo] == key) ? lo : UINT64_MAX;
}

// findassemblystart.cpp:12-23
int findPosOfM(const char *seq) {
    int pos = 0;
    while (seq[pos] != '\0') {
        if (seq[pos] == 'M') return pos;
        pos++;
    }
    return -1;
}

struct PositionOfM {     // findassemblystart.cpp:25-33
    uint64_t id;
    int mPos;
    bool hasM;
    bool hasStopM;
};

void emitDb(const std::vector<std::string> &entries, const or_seqdb *db, char **out_data, uint64_t **out_offsets,
            uint32_t **out_lens, uint32_t **out_keys, uint64_t *out_n, uint64_t *out_bytes,
            const std::vector<uint64_t> &ids) {
    std::string all;
    const uint64_t n = ids.size();
    uint64_t *offs = (uint64_t *) malloc(sizeof(uint64_t) * (n + 1));
    uint32_t *lens = (uint32_t *) malloc(sizeof(uint32_t) * (n + 1));
    uint32_t *keys = (uint32_t *) malloc(sizeof(uint32_t) * (n + 1));
    for (uint64_t w = 0; w < n; w++) {
        offs[w] = all.size();
        all.append(entries[w]);
        lens[w] = (uint32_t) entries[w].size();
        keys[w] = db->keys[ids[w]];
    }
    *out_data = (char *) malloc(all.size() + 1);
    memcpy(*out_data, all.data(), all.size());
    *out_offsets = offs; *out_lens = lens; *out_keys = keys; *out_n = n; *out_bytes = all.size();
}

}  // namespace

/* findassemblystart (src/assembler/findassemblystart.cpp:35-176).  alns = the alignment DB of the same sequence DB
 * (query == target DB), ordered by query key, lines in DB order.  add_stop[i] (by sequence index) = the position the
 * new '*' is put in front of, or -1.  Output entries: unchanged, or "*" + residues[mPos..] + "\n\0" (:150-160). */
extern "C" int or_findstart(const or_seqdb *db, const or_aln *alns, uint64_t n_alns,
                            char **out_data, uint64_t **out_offsets, uint32_t **out_lens, uint32_t **out_keys,
                            uint64_t *out_n, uint64_t *out_bytes, int32_t **add_stop) {
    std::vector<int> addStopAtPosition(db->n, -1);
    const float threshold = 0.2;
    std::vector<PositionOfM> stopPositions;
    uint64_t ap = 0;
    while (ap < n_alns) {
        const uint32_t queryKey = alns[ap].query;
        uint64_t ae = ap;
        while (ae < n_alns && alns[ae].query == queryKey) ae++;
        const uint64_t qId = idOfKey(db, queryKey);
        if (qId == UINT64_MAX) return -1;
        const char *querySeqData = db->data + db->offsets[qId];
        const int queryPosOfM = findPosOfM(querySeqData);
        if (queryPosOfM == -1) { ap = ae; continue; }
        bool qStopM = false;
        if (queryPosOfM > 0) qStopM = querySeqData[queryPosOfM - 1] == '*';
        stopPositions.clear();
        stopPositions.push_back(PositionOfM{qId, queryPosOfM, true, qStopM});
        for (uint64_t a = ap; a < ae; a++) {
            const uint64_t edgeId = idOfKey(db, alns[a].target);
            if (edgeId == UINT64_MAX) return -1;
            if (edgeId == qId) continue;
            const or_aln &res = alns[a];
            const char *dbSeqData = db->data + db->offsets[edgeId];
            int posOfM = -1;
            bool hasM = false, hasStopM = false;
            if (res.q_start >= queryPosOfM && queryPosOfM <= res.q_end) {     // :103 (sic)
                int queryMoffset = queryPosOfM - res.q_start;
                int dbMPos = res.db_start + queryMoffset;
                posOfM = dbMPos;
                hasM = dbMPos >= 0 && (dbSeqData[dbMPos] == 'M');
                if (dbMPos > 0 && hasM) hasStopM = dbSeqData[dbMPos - 1] == '*';
            }
            stopPositions.push_back(PositionOfM{edgeId, posOfM, hasM, hasStopM});
        }
        int stopMCount = 0;
        for (size_t i = 0; i < stopPositions.size(); i++) stopMCount += stopPositions[i].hasStopM;
        if (stopPositions.size() > 1) {
            const float frequency = static_cast<float>(stopMCount) / static_cast<float>(stopPositions.size());
            if (frequency >= threshold) {
                for (size_t i = 0; i < stopPositions.size(); i++) {
                    int &t = addStopAtPosition[stopPositions[i].id];
                    if (t < stopPositions[i].mPos) t = stopPositions[i].mPos;      // the CAS loop of :125-131 = atomic max
                }
            }
        }
        ap = ae;
    }
    std::vector<std::string> entries(db->n);
    std::vector<uint64_t> ids(db->n);
    for (uint64_t id = 0; id < db->n; id++) {
        ids[id] = id;
        const char *q = db->data + db->offsets[id];
        const int mPos = addStopAtPosition[id];
        if (mPos == -1) {
            entries[id].assign(q, db->lens[id] - 1);      // getEntryLen - 1 bytes, writeData appends '\0'
        } else {
            entries[id] = "*";
            entries[id].append(q + mPos);                 // up to the entry's '\0': residues + '\n'
        }
        entries[id].push_back('\0');
    }
    emitDb(entries, db, out_data, out_offsets, out_lens, out_keys, out_n, out_bytes, ids);
    int32_t *as = (int32_t *) malloc(sizeof(int32_t) * (db->n + 1));
    for (uint64_t id = 0; id < db->n; id++) as[id] = addStopAtPosition[id];
    *add_stop = as;
    return 0;
}

/* cyclecheck (src/assembler/cyclecheck.cpp:71-274), k = 22 (setCycleCheckDefaults :26-29; -k is not a cyclecheck flag).
 * split[i] = splitDiagonal of sequence i (0 = not reported).  The output DB holds the reported sequences only:
 * chop_cycle ? residues[0 .. splitDiagonal) + "\n\0" : the entry unchanged (:249-259). */
extern "C" int or_cyclecheck(const or_seqdb *db, int max_seq_len, int kmer_size, uint32_t *split) {
    if (db->dbtype != 1) return -1;
    // NucleotideMatrix aa2num (mm/commons/NucleotideMatrix.cpp:17-61), dumped from the reference's own object
    const unsigned char *aa2num = OR_NT_AA2NUM;
    const size_t kmerSize = (size_t) kmer_size;
    // Indexer(alphabetSize - 1 = 4, kmerSize): powers[i] = 4^i (mm/prefiltering/Indexer.cpp:4-21)
    std::vector<uint64_t> powers(kmerSize);
    { uint64_t pw = 1; for (size_t i = 0; i < kmerSize; i++) { powers[i] = pw; pw *= 4; } }
    struct kmerSeqPos { uint64_t kmer; unsigned int pos; };
    auto cmp = [](const kmerSeqPos &a, const kmerSeqPos &b) { if (a.kmer != b.kmer) return a.kmer < b.kmer; return a.pos < b.pos; };
    std::vector<kmerSeqPos> frontKmers, middleKmers, backKmers;
    std::vector<unsigned int> diagHits;
    std::vector<unsigned char> num;
    for (uint64_t id = 0; id < db->n; id++) {
        split[id] = 0;
        const char *nuclSeq = db->data + db->offsets[id];
        const unsigned int seqLen = db->lens[id] - 2;
        if (seqLen >= (unsigned int) max_seq_len) continue;                      // :107-112
        // Sequence::mapSequence (mm/commons/Sequence.cpp:476-489): stops at '\0' / '\n'
        num.assign(seqLen, 0);
        unsigned int L = 0;
        while (L < seqLen && nuclSeq[L] != '\0' && nuclSeq[L] != '\n') { num[L] = aa2num[(unsigned char) nuclSeq[L]]; L++; }
        frontKmers.clear(); middleKmers.clear(); backKmers.clear();
        const unsigned int thirdSeqLen = seqLen / 3;
        // Sequence::hasNextKmer / nextKmer (mm/commons/Sequence.h:98-114): currItPos starts at -1; note that `pos` is read
        // BEFORE nextKmer advances it (:124), i.e. it is the window position minus one, 0xFFFFFFFF for the first window.
        int currItPos = -1;
        while ((size_t) (currItPos + 1) + kmerSize <= (size_t) L) {
            unsigned int pos = (unsigned int) currItPos;
            currItPos++;
            uint64_t kmerIdx = 0;
            for (size_t i = 0; i < kmerSize; i++) kmerIdx += num[currItPos + i] * powers[i];
            kmerSeqPos e; e.kmer = kmerIdx; e.pos = (unsigned int) currItPos;
            if (pos < thirdSeqLen + 1) frontKmers.push_back(e);
            else if (pos < 2 * thirdSeqLen + 1) middleKmers.push_back(e);
            else backKmers.push_back(e);
        }
        std::sort(frontKmers.begin(), frontKmers.end(), cmp);
        std::sort(middleKmers.begin(), middleKmers.end(), cmp);
        std::sort(backKmers.begin(), backKmers.end(), cmp);
        const unsigned int frontKmersCount = frontKmers.size(), middleKmersCount = middleKmers.size(), backKmersCount = backKmers.size();
        unsigned int kmermatches = 0;
        diagHits.assign(2 * (size_t) thirdSeqLen + 1, 0);
        unsigned int idx = 0, jdx = 0, kdx = 0;
        while (idx < frontKmersCount && (jdx < backKmersCount || kdx < middleKmersCount)) {      // :150-186
            uint64_t kmerIdx = frontKmers[idx].kmer;
            unsigned int pos = frontKmers[idx].pos;
            while (jdx < backKmersCount && backKmers[jdx].kmer < kmerIdx) jdx++;
            while (kdx < middleKmersCount && middleKmers[kdx].kmer < kmerIdx) kdx++;
            while (jdx < backKmersCount && kmerIdx == backKmers[jdx].kmer) {
                int diag = backKmers[jdx].pos - pos;
                if (diag >= static_cast<int>(seqLen / 3)) { diagHits[diag - seqLen / 3]++; kmermatches++; }
                jdx++;
            }
            while (kdx < middleKmersCount && kmerIdx == middleKmers[kdx].kmer) {
                int diag = middleKmers[kdx].pos - pos;
                if (diag >= static_cast<int>(seqLen / 3)) { diagHits[diag - seqLen / 3]++; kmermatches++; }
                kdx++;
            }
            idx++;
            while (idx < frontKmersCount && kmerIdx == frontKmers[idx].kmer) idx++;
        }
        jdx = 0, kdx = 0;
        while (kdx < middleKmersCount && jdx < backKmersCount) {                                   // :189-213
            if (middleKmers[kdx].kmer < backKmers[jdx].kmer) kdx++;
            else if (middleKmers[kdx].kmer > backKmers[jdx].kmer) jdx++;
            else {
                uint64_t kmerIdx = middleKmers[kdx].kmer;
                unsigned int pos = middleKmers[kdx].pos;
                while (jdx < backKmersCount && kmerIdx == backKmers[jdx].kmer) {
                    int diag = backKmers[jdx].pos - pos;
                    if (diag >= static_cast<int>(seqLen / 3)) { diagHits[diag - seqLen / 3]++; kmermatches++; }
                    jdx++;
                }
                while (kdx < middleKmersCount && kmerIdx == middleKmers[kdx].kmer) kdx++;
            }
        }
        unsigned int splitDiagonal = 0;
        if (kmermatches > 0) {                                                                     // :238-262
            for (unsigned int d = 0; d < 2 * thirdSeqLen; d++) {
                if (diagHits[d] != 0) {
                    unsigned int diag = d + thirdSeqLen;
                    unsigned int diaglen = seqLen - diag;
                    unsigned int gapwindow = diaglen * 0.01;
                    unsigned int lower = std::max(0, static_cast<int>(d - gapwindow));
                    unsigned int upper = std::min(d + gapwindow, 2 * thirdSeqLen);
                    unsigned int diagbandHits = 0;
                    for (size_t i = lower; i <= upper; i++)
                        if (diagHits[i] <= diagHits[d]) diagbandHits += diagHits[i];
                    float diagbandHitRate = static_cast<float>(diagbandHits) / (diaglen - kmerSize + 1);
                    if (diagbandHitRate > 0.2) { splitDiagonal = diag; break; }
                }
            }
        }
        split[id] = splitDiagonal;
    }
    return 0;
}

// =================================================================================================================
// extractorfs (+ translatenucs --add-orf-stop 1): the six-frame ORF pipeline upstream of the first iteration
// (data/assemble.sh:41-77; mm/util/extractorfs.cpp:20-159, mm/commons/Orf.cpp:127-330, mm/commons/TranslateNucl.h,
//  mm/util/translatenucs.cpp:14-128).  Translation table 1 (canonical) only.
// =================================================================================================================
namespace {

const char *IUPAC_RC =           // Orf::iupacReverseComplementTable (Orf.cpp:48-52)
    "................................................................"
    ".TVGH..CD..M.KN...YSAABW.R.......tvgh..cd..m.kn...ysaabw.r......"
    "................................................................"
    "................................................................";

struct Translate {               // TranslateNucl(CANONICAL): sm_BaseToIdx + m_AminoAcid (TranslateNucl.h:333-470)
    int baseToIdx[256];
    char aminoAcid[4097];
    Translate() {
        static const char charToBase[17] = "-ACMGRSVTWYHKDBN";
        for (int i = 0; i < 256; i++) baseToIdx[i] = 0;
        for (int i = 0; i <= 15; i++) { baseToIdx[(int) charToBase[i]] = i; baseToIdx[(int) (unsigned char) tolower(charToBase[i])] = i; }
        baseToIdx[(int) 'U'] = 8; baseToIdx[(int) 'u'] = 8; baseToIdx[(int) 'X'] = 15; baseToIdx[(int) 'x'] = 15;
        for (int i = 0; i <= 15; i++) baseToIdx[i] = i;
        const std::string ncbieaa = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
        static const int expansions[4] = {1, 2, 4, 8};                     // A C G T
        static const int codonIdx[9] = {0, 2, 1, 0, 3, 0, 0, 0, 0};        // T = 0, C = 1, A = 2, G = 3
        for (int i = 0; i <= 4096; i++) aminoAcid[i] = 'X';
        int st = 1;
        for (int i = 0; i <= 15; i++) for (int j = 0; j <= 15; j++) for (int k = 0; k <= 15; k++, st++) {
            char aa = '\0';
            bool go_on = true;
            for (int p = 0; p < 4 && go_on; p++) { int x = expansions[p]; if ((x & i) == 0) continue;
                for (int q = 0; q < 4 && go_on; q++) { int y = expansions[q]; if ((y & j) == 0) continue;
                    for (int r = 0; r < 4 && go_on; r++) { int z = expansions[r]; if ((z & k) == 0) continue;
                        char ch = ncbieaa[16 * codonIdx[x] + 4 * codonIdx[y] + codonIdx[z]];
                        if (aa == '\0') aa = ch;
                        else if (aa != ch) {
                            if ((aa == 'B' || aa == 'D' || aa == 'N') && (ch == 'D' || ch == 'N')) aa = 'B';
                            else if ((aa == 'Z' || aa == 'E' || aa == 'Q') && (ch == 'E' || ch == 'Q')) aa = 'Z';
                            else if ((aa == 'J' || aa == 'I' || aa == 'L') && (ch == 'I' || ch == 'L')) aa = 'J';
                            else aa = 'X';
                        }
                        // the reference leaves the loop once both the amino acid and the start flag are 'X'; the amino acid
                        // cannot change after it became 'X', so stopping here gives the same table
                        if (aa == 'X') go_on = false;
                    } } }
            if (aa != '\0') aminoAcid[st] = aa;
        }
    }
    // TranslateNucl::translate (TranslateNucl.h:488-503); the codon state only depends on the last three letters
    void translate(char *aa, const char *nucl, int L) const {
        for (int i = 0; i < L; i += 3) {
            bool lower = false;
            int idx[3];
            for (int k = 0; k < 3; k++) { lower |= islower((unsigned char) nucl[i + k]) != 0; idx[k] = baseToIdx[(unsigned char) nucl[i + k]]; }
            char residue = aminoAcid[256 * idx[0] + 16 * idx[1] + idx[2] + 1];
            aa[i / 3] = lower ? (char) tolower(residue) : residue;
        }
    }
};

struct OrfLoc { size_t from, to; bool incStart, incEnd; int strand; };

inline bool codonIs(const char *c, const char *s) { return c[0] == s[0] && c[1] == s[1] && c[2] == s[2]; }

// Orf::findForward (Orf.cpp:192-330)
void findForward(const char *sequence, size_t sequenceLength, std::vector<OrfLoc> &result, size_t minLength, size_t maxLength,
                 size_t maxGaps, unsigned frames, unsigned startMode, int strand, bool allStarts) {
    const unsigned frameLookup[3] = {1u, 2u, 4u};
    bool isInsideOrf[3] = {true, true, true}, hasStartCodon[3] = {false, false, false};
    size_t countGaps[3] = {0, 0, 0}, countLength[3] = {0, 0, 0}, from[3] = {0, 1, 2};
    auto isIncomplete = [](const char *c) { return c[0] == CHAR_MAX || c[1] == CHAR_MAX || c[2] == CHAR_MAX; };
    auto isStart = [&](const char *c) { return codonIs(c, "ATG") || (allStarts && (codonIs(c, "TTG") || codonIs(c, "CTG"))); };
    auto isStop = [](const char *c) { return codonIs(c, "TAA") || codonIs(c, "TAG") || codonIs(c, "TGA"); };
    auto isGapOrN = [](const char *c) {
        return c[0] == 'N' || IUPAC_RC[(unsigned char) c[0]] == '.' || c[1] == 'N' || IUPAC_RC[(unsigned char) c[1]] == '.'
            || c[2] == 'N' || IUPAC_RC[(unsigned char) c[2]] == '.';
    };
    for (size_t i = 0; i < sequenceLength - 2; i += 3) {
        for (size_t position = i; position < i + 3; position++) {
            char codon[3];
            for (int k = 0; k < 3; k++)
                codon[k] = sequence[position + k] == CHAR_MAX ? (char) CHAR_MAX : (char) (sequence[position + k] & static_cast<unsigned char>(~0x20));
            size_t frame = position % 3;
            if (!(frames & frameLookup[frame])) continue;
            bool thisIncomplete = isIncomplete(codon);
            bool isLast = !thisIncomplete && isIncomplete(sequence + position + 3);
            bool shouldStart;
            if (startMode == 0) shouldStart = isInsideOrf[frame] == false && isStart(codon);
            else if (startMode == 1) shouldStart = isInsideOrf[frame] == false;
            else shouldStart = isStart(codon);
            if (shouldStart) { isInsideOrf[frame] = true; hasStartCodon[frame] = true; from[frame] = position; countGaps[frame] = 0; countLength[frame] = 0; }
            const bool stop = isStop(codon);
            if (isInsideOrf[frame]) {
                if (!stop) countLength[frame]++;
                if (isGapOrN(codon)) countGaps[frame]++;
            }
            if (isInsideOrf[frame] && (stop || isLast)) {
                isInsideOrf[frame] = false;
                if (countLength[frame] == 0 && stop) continue;
                size_t to = position + ((isLast && stop == false) ? 2 : -1);
                if (countGaps[frame] > maxGaps || countLength[frame] > maxLength || countLength[frame] < minLength) continue;
                result.push_back(OrfLoc{from[frame], to, !hasStartCodon[frame], !stop, strand});
            }
        }
    }
}

}  // namespace

/* extractorfs, optionally fused with translatenucs --add-orf-stop 1 (translate != 0).  Output: fragments in (read index,
 * emission order) with keys 0..n-1 (= the renumbered DB, DBWriter::createRenumberedDB) and, per fragment, the ORF header
 * fields {read key, fromPos, toPos, incompleteStart | incompleteEnd << 1} (Orf::writeOrfHeader, Orf.cpp:445-462). */
extern "C" int or_extractorfs(const or_seqdb *db, const or_orf_params *p, int translate,
                              char **out_data, uint64_t **out_offsets, uint32_t **out_lens, uint32_t **out_keys,
                              uint64_t *out_n, uint64_t *out_bytes, uint32_t **orf_info /* 4 x n */) {
    if (p->translation_table != 1) return -1;
    if (p->orf_start_mode == 1 && p->contig_start_mode < 2) return -2;     // extractorfs.cpp:38-41
    static const Translate T;
    std::string all;
    std::vector<uint64_t> offs;
    std::vector<uint32_t> lens, info;
    std::vector<char> sequence, reverseComplement, aa;
    std::vector<OrfLoc> res;
    for (uint64_t i = 0; i < db->n; i++) {
        const char *data = db->data + db->offsets[i];
        const size_t L = db->lens[i] - 2;
        if (L < 3) continue;                                               // Orf::setSequence (Orf.cpp:127-131)
        sequence.assign(L + 32, (char) CHAR_MAX);
        reverseComplement.assign(L + 32, (char) CHAR_MAX);
        for (size_t k = 0; k < L; k++) sequence[k] = (data[k] == 'u') ? 't' : data[k];     // sic: the 'U' branch is overwritten (:144-147)
        for (size_t k = 0; k < L; k++) {
            char c = IUPAC_RC[(unsigned char) sequence[L - k - 1]];
            reverseComplement[k] = (c == '.') ? 'N' : c;
        }
        res.clear();
        if (p->forward_frames != 0) findForward(sequence.data(), L, res, p->min_length, p->max_length, p->max_gaps, p->forward_frames, p->orf_start_mode, +1, p->use_all_table_starts != 0);
        if (p->reverse_frames != 0) findForward(reverseComplement.data(), L, res, p->min_length, p->max_length, p->max_gaps, p->reverse_frames, p->orf_start_mode, -1, p->use_all_table_starts != 0);
        for (const OrfLoc &loc : res) {
            if (p->contig_start_mode < 2 && ((int) loc.incStart == p->contig_start_mode)) continue;
            if (p->contig_end_mode < 2 && ((int) loc.incEnd == p->contig_end_mode)) continue;
            const char *s = (loc.strand > 0 ? sequence.data() : reverseComplement.data()) + loc.from;
            const size_t len = loc.to - loc.from + 1;
            size_t fromPos = loc.from, toPos = loc.to;
            if (loc.strand < 0) { fromPos = (L - 1) - loc.from; toPos = (L - 1) - loc.to; }
            offs.push_back(all.size());
            if (!translate) {
                all.append(s, len); all.push_back('\n'); all.push_back('\0');
                lens.push_back((uint32_t) len + 2);
            } else {
                // translatenucs.cpp:55-118 on the entry "s[0..len) \n": len is a multiple of three
                const bool addStopAtStart = !loc.incStart;
                bool addStopAtEnd = !loc.incEnd;
                aa.assign(len / 3 + 4, 0);
                char *w = aa.data();
                if (addStopAtStart) *w++ = '*';
                T.translate(w, s, (int) len);
                if (addStopAtEnd && w[len / 3 - 1] != '*') { w[len / 3] = '*'; w[len / 3 + 1] = '\n'; }
                else { addStopAtEnd = false; w[len / 3] = '\n'; }
                const size_t n = len / 3 + 1 + addStopAtStart + addStopAtEnd;
                all.append(aa.data(), n); all.push_back('\0');
                lens.push_back((uint32_t) n + 1);
            }
            info.push_back(db->keys[i]); info.push_back((uint32_t) fromPos); info.push_back((uint32_t) toPos);
            info.push_back((uint32_t) loc.incStart | ((uint32_t) loc.incEnd << 1));
        }
    }
    const uint64_t n = offs.size();
    *out_data = (char *) malloc(all.size() + 1);
    memcpy(*out_data, all.data(), all.size());
    *out_offsets = (uint64_t *) malloc(sizeof(uint64_t) * (n + 1));
    *out_lens = (uint32_t *) malloc(sizeof(uint32_t) * (n + 1));
    *out_keys = (uint32_t *) malloc(sizeof(uint32_t) * (n + 1));
    *orf_info = (uint32_t *) malloc(sizeof(uint32_t) * (4 * n + 4));
    for (uint64_t k = 0; k < n; k++) { (*out_offsets)[k] = offs[k]; (*out_lens)[k] = lens[k]; (*out_keys)[k] = (uint32_t) k; }
    if (n) memcpy(*orf_info, info.data(), sizeof(uint32_t) * 4 * n);
    *out_n = n; *out_bytes = all.size();
    return 0;
}

/* translatenucs on any nucleotide DB (mm/util/translatenucs.cpp:14-128).  flags (may be NULL = --add-orf-stop 0): per
 * sequence bit 0 = addStopAtStart, bit 1 = addStopAtEnd (from the ORF header, :58-63).  Entries shorter than a codon are
 * dropped; keys are kept. */
extern "C" int or_translatenucs(const or_seqdb *db, const uint8_t *flags, int max_seq_len,
                                char **out_data, uint64_t **out_offsets, uint32_t **out_lens, uint32_t **out_keys,
                                uint64_t *out_n, uint64_t *out_bytes) {
    static const Translate T;
    std::vector<std::string> entries;
    std::vector<uint64_t> ids;
    std::vector<char> aa;
    for (uint64_t i = 0; i < db->n; i++) {
        const char *data = db->data + db->offsets[i];
        if (*data == '\0') continue;
        bool addStopAtStart = flags ? (flags[i] & 1) != 0 : false;
        bool addStopAtEnd = flags ? (flags[i] & 2) != 0 : false;
        size_t length = db->lens[i] - 1;
        if ((data[length] != '\n' && length % 3 != 0) && (data[length - 1] == '\n' && (length - 1) % 3 != 0)) length = length - (length % 3);
        if (length < 3) continue;
        if (length > (size_t) (3 * max_seq_len)) length = 3 * max_seq_len;
        aa.assign(length / 3 + 8, 0);
        char *writeAA = aa.data();
        if (addStopAtStart) { aa[0] = '*'; writeAA = aa.data() + 1; }
        // the reference translates ceil(length / 3) codons and then overwrites position length / 3 with the terminator;
        // the last (partial) codon may read past the entry, its result never survives: translate only the full ones
        T.translate(writeAA, data, (int) (length / 3) * 3);
        if (addStopAtEnd && writeAA[(length / 3) - 1] != '*') { writeAA[length / 3] = '*'; writeAA[length / 3 + 1] = '\n'; }
        else { addStopAtEnd = false; writeAA[length / 3] = '\n'; }
        std::string e(aa.data(), (length / 3) + 1 + addStopAtStart + addStopAtEnd);
        e.push_back('\0');
        entries.push_back(e);
        ids.push_back(i);
    }
    emitDb(entries, db, out_data, out_offsets, out_lens, out_keys, out_n, out_bytes, ids);
    return 0;
}
