// oracle.cpp -- CPU restatement of the reference's assemble-iteration hot path.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Scalar, single-threaded, written for clarity and for
// bug-compatibility with the reference, not for speed.  Every function cites the reference code it
// follows ("mm/" = /root/reference/lib/mmseqs/src/).  Pinned against the unmodified reference binary
// by tests/test_oracle_vs_reference.py (fixtures under tests/golden/).
#include "oracle.h"
#include "oracle_tables.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <string>
#include <vector>

namespace {

const uint64_t U64MAX = ~0ULL;
inline uint64_t bitSet63(uint64_t v) { return v | (1ULL << 63); }
inline uint64_t bitClear63(uint64_t v) { return v & ~(1ULL << 63); }
inline bool bitCheck63(uint64_t v) { return (v >> 63) & 1; }

// ---------------------------------------------------------------------------------------------
// XXH64 specialised to an 8-byte input (lib/mmseqs/lib/xxhash/xxhash.h, XXH64_endian_align with
// len == 8: no stripe loop, one 8-byte round in the tail, then the avalanche).
// ---------------------------------------------------------------------------------------------
const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL,
               P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
uint64_t xxh64_u64(uint64_t v, uint64_t seed) {
    uint64_t h = seed + P5 + 8;
    uint64_t k = rotl(v * P2, 31) * P1;
    h ^= k;
    h = rotl(h, 27) * P1 + P4;
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

// Util::revComplement (mm/commons/Util.cpp:601-638): 2-bit code A0 C1 T2 G3, complement = xor 2.
uint64_t revComplement(uint64_t kmer, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) {
        r = (r << 2) | ((kmer & 3) ^ 2);
        kmer >>= 2;
    }
    return r;
}

struct Seq {
    const char *s;
    int len;      // getSeqLen = entry len - 2 (mm/commons/DBReader.h:192-213)
    uint32_t key;
};

inline Seq getSeq(const or_seqdb *db, uint64_t i) {
    Seq q;
    q.s = db->data + db->offsets[i];
    q.len = (int) db->lens[i] - 2;
    if (q.len < 0) q.len = 0;
    q.key = db->keys[i];
    return q;
}

// DBReader::getId: binary search for a key in the key-sorted index (mm/commons/DBReader.cpp).
uint64_t findId(const or_seqdb *db, uint32_t key) {
    const uint32_t *b = db->keys, *e = db->keys + db->n;
    const uint32_t *it = std::lower_bound(b, e, key);
    if (it == e || *it != key) return U64MAX;
    return (uint64_t) (it - b);
}

// ---------------------------------------------------------------------------------------------
// kmermatcher
// ---------------------------------------------------------------------------------------------
struct SeqPos {      // SequencePosition (mm/linclust/kmermatcher.h:10-46)
    unsigned short score;
    uint64_t kmer;
    unsigned int pos;
};

bool wideRecords(const or_seqdb *db) {   // kmermatcher.cpp:797-802: T = short iff maxSeqLen < SHRT_MAX
    int mx = 0;
    for (uint64_t i = 0; i < db->n; i++) mx = std::max(mx, (int) db->lens[i] - 2);
    return !(mx < SHRT_MAX);
}

// fillKmerPositionArray for one sequence (kmermatcher.cpp:126-347).
void extractOne(const or_seqdb *db, uint64_t idx, const or_km_params *p, bool nt,
                const unsigned char *aa2num, int xCode, std::vector<or_kmer_rec> &out,
                std::vector<unsigned char> &num, std::vector<SeqPos> &kmers,
                std::vector<unsigned short> &scoreDist, std::vector<unsigned int> &hier) {
    Seq q = getSeq(db, idx);
    const int k = p->kmer_size;
    // Sequence::mapSequence (mm/commons/Sequence.cpp:476-489): stop at '\0' / '\n'
    num.clear();
    int L = 0;
    while (L < q.len && q.s[L] != '\0' && q.s[L] != '\n') {
        num.push_back(aa2num[(unsigned char) q.s[L]]);
        L++;
    }
    // whole-sequence hash (kmermatcher.cpp:133-138; Util::hash poly-31 over codes, Util.h:336-345)
    uint64_t h = 0;
    for (int i = 0; i < L; i++) h = h * 31 + num[i];
    uint64_t seqHash = xxh64_u64(h, (uint64_t) p->hash_shift);

    std::fill(scoreDist.begin(), scoreDist.end(), 0);
    std::fill(hier.begin(), hier.end(), 0);
    kmers.clear();
    // aa index base = alphabetSize-1 (kmermatcher.cpp:113; Indexer.cpp:4-21)
    const uint64_t base = nt ? 4 : (uint64_t) (p->alph_size - 1);
    for (int pos = 0; pos + k <= L; pos++) {            // hasNextKmer/nextKmer (Sequence.h:98-420)
        bool hasX = false;
        for (int j = 0; j < k; j++) hasX |= (num[pos + j] == xCode);
        if (hasX) continue;                             // kmerContainsX (kmermatcher.cpp:146-148)
        SeqPos sp;
        if (nt) {
            uint64_t kmerIdx = 0;                       // Indexer::computeKmerIdx (Indexer.h:124-131)
            for (int j = 0; j < k; j++) kmerIdx = (kmerIdx << 2) | num[pos + j];
            uint64_t rev = revComplement(kmerIdx, k);
            if (rev == kmerIdx) continue;               // kmermatcher.cpp:156-158
            bool pickRev = rev < kmerIdx;
            kmerIdx = pickRev ? rev : kmerIdx;
            sp.score = (unsigned short) xxh64_u64(kmerIdx, (uint64_t) p->hash_shift);
            sp.kmer = pickRev ? bitClear63(kmerIdx) : bitSet63(kmerIdx);   // :181
            sp.pos = pickRev ? (unsigned int) (L - pos - k) : (unsigned int) pos;   // :184
        } else {
            uint64_t kmerIdx = 0, pw = 1;               // Indexer::int2index (Indexer.h:20-83)
            for (int j = 0; j < k; j++) { kmerIdx += (uint64_t) num[pos + j] * pw; pw *= base; }
            sp.kmer = kmerIdx;
            sp.pos = (unsigned int) pos;
            sp.score = (unsigned short) xxh64_u64(kmerIdx, (uint64_t) p->hash_shift);
        }
        scoreDist[sp.score]++;
        hier[sp.score >> 9]++;
        kmers.push_back(sp);
    }
    size_t seqKmerCount = kmers.size();
    // kmermatcher.cpp:223 -- float arithmetic then size_t cast
    size_t kmerConsidered = std::min(static_cast<size_t>(p->kmers_per_seq - 1 + (p->kmers_per_seq_scale * L)), seqKmerCount);
    unsigned int threshold = 0;
    size_t kmerInBins = 0;
    if (seqKmerCount > 0) {                             // :227-237
        size_t ht = 0;
        for (ht = 0; ht < 128 && kmerInBins < kmerConsidered; ht++) kmerInBins += hier[ht];
        ht -= (ht > 0) ? 1 : 0;
        kmerInBins -= hier[ht];
        for (threshold = ht * 512; threshold <= USHRT_MAX && kmerInBins < kmerConsidered; threshold++)
            kmerInBins += scoreDist[threshold];
    }
    int tooMuch = (int) (kmerInBins - kmerConsidered);  // :238
    // sequence-identity record (:241-246)
    if ((unsigned short) seqHash >= p->hash_start && (unsigned short) seqHash <= p->hash_end) {
        or_kmer_rec r; r.kmer = seqHash; r.id = q.key; r.pos = 0; r.seq_len = L;
        out.push_back(r);
    }
    if (p->ignore_multi_kmer) {                         // :266-272 (SequencePosition::compareByScore[Reverse])
        std::sort(kmers.begin(), kmers.end(), [nt](const SeqPos &a, const SeqPos &b) {
            if (a.score != b.score) return a.score < b.score;
            uint64_t ka = nt ? bitSet63(a.kmer) : a.kmer, kb = nt ? bitSet63(b.kmer) : b.kmer;
            if (ka != kb) return ka < kb;
            return a.pos < b.pos;
        });
    }
    size_t selected = 0;
    for (size_t i = 0; i < seqKmerCount && selected < kmerConsidered; i++) {   // :274-347
        if (p->ignore_multi_kmer) {
            uint64_t kmer = kmers[i].kmer;
            if (nt) kmer = bitSet63(kmer);
            if (i + 1 < seqKmerCount) {
                uint64_t next = kmers[i + 1].kmer;
                if (nt) next = bitSet63(next);
                if (kmer == next) {
                    while (kmer == next && i < seqKmerCount) {
                        i++;
                        if (i >= seqKmerCount) break;
                        next = kmers[i].kmer;
                        if (nt) next = bitSet63(next);
                    }
                }
            }
            if (i >= seqKmerCount) break;
        }
        if (kmers[i].score < threshold) {
            if (kmers[i].score == (threshold - 1) && tooMuch) {
                tooMuch--;
                threshold -= (tooMuch == 0) ? 1 : 0;
            }
            selected++;
            if (kmers[i].score >= p->hash_start && kmers[i].score <= p->hash_end) {
                or_kmer_rec r; r.kmer = kmers[i].kmer; r.id = q.key; r.pos = (int32_t) kmers[i].pos; r.seq_len = L;
                out.push_back(r);
            }
        }
    }
}

void extractAll(const or_seqdb *db, const or_km_params *p, std::vector<or_kmer_rec> &out) {
    const bool nt = db->dbtype == 1;
    const unsigned char *aa2num = nt ? OR_NT_AA2NUM : (p->alph_size == 21 ? OR_AA_AA2NUM : OR_RED_AA2NUM);
    const int xCode = nt ? 4 : (p->alph_size - 1);
    std::vector<unsigned char> num;
    std::vector<SeqPos> kmers;
    std::vector<unsigned short> scoreDist(65536);
    std::vector<unsigned int> hier(128);
    for (uint64_t i = 0; i < db->n; i++)
        extractOne(db, i, p, nt, aa2num, xCode, out, num, kmers, scoreDist, hier);
}

// Util::canBeCovered (mm/commons/Util.cpp:533-551)
bool canBeCovered(float covThr, int covMode, float q, float t) {
    switch (covMode) {
        case 0: return (q / t >= covThr) && (t / q >= covThr);
        case 1: return (t / q) >= covThr;
        case 2: return (q / t) >= covThr;
        case 3: return (t / q) >= covThr && (t / q) <= 1.0;
        case 4: return (q / t) >= covThr && (q / t) <= 1.0;
        case 5: return (std::min(t, q) / std::max(t, q)) >= covThr;
        default: return true;
    }
}
bool hasCoverage(float covThr, int covMode, float qc, float tc) {   // Util.cpp:553-567
    switch (covMode) {
        case 0: return qc >= covThr && tc >= covThr;
        case 1: return qc >= covThr;
        case 2: return tc >= covThr;
        default: return true;
    }
}

}  // namespace

extern "C" uint64_t or_hash_u64(uint64_t v, uint64_t seed) { return xxh64_u64(v, seed); }

extern "C" int or_extract_kmers(const or_seqdb *db, const or_km_params *p, or_kmer_rec **out, uint64_t *n_out) {
    std::vector<or_kmer_rec> v;
    extractAll(db, p, v);
    *n_out = v.size();
    *out = (or_kmer_rec *) malloc(sizeof(or_kmer_rec) * (v.size() + 1));
    if (!v.empty()) memcpy(*out, v.data(), sizeof(or_kmer_rec) * v.size());
    return 0;
}

extern "C" int or_kmermatch(const or_seqdb *db, const or_km_params *p, or_hit **out, uint64_t *n_out) {
    const bool nt = db->dbtype == 1;
    const bool wide = wideRecords(db);
    // computeKmerCount (kmermatcher.cpp:576-585) -> array of totalKmers+1(+1) records preset to 0xFF
    // (initKmerPositionMemory :40-54; kmermatcherInner :617-622 with enough memory => one split).
    size_t totalKmers = 0;
    for (uint64_t i = 0; i < db->n; i++) {
        int seqLen = (int) db->lens[i] - 2;
        int adj = std::max(1, seqLen - p->kmer_size + 2);
        totalKmers += std::min(adj, static_cast<int>(p->kmers_per_seq + (p->kmers_per_seq_scale * seqLen)));
    }
    size_t totalKmersPerSplit = std::max((size_t) 1025, totalKmers + 1);
    std::vector<or_kmer_rec> a;
    extractAll(db, p, a);
    const size_t elementsToSort = a.size();
    or_kmer_rec sentinel; memset(&sentinel, 0xFF, sizeof(sentinel));
    a.resize(totalKmersPerSplit + 1, sentinel);
    auto tshort = [wide](int v) -> int32_t { return wide ? v : (int32_t) (short) v; };

    // sort #1 (kmermatcher.cpp:408-412; comparators kmermatcher.h:56-96)
    std::sort(a.begin(), a.begin() + elementsToSort, [nt](const or_kmer_rec &x, const or_kmer_rec &y) {
        uint64_t kx = nt ? bitSet63(x.kmer) : x.kmer, ky = nt ? bitSet63(y.kmer) : y.kmer;
        if (kx != ky) return kx < ky;
        if (x.seq_len != y.seq_len) return x.seq_len > y.seq_len;
        if (x.id != y.id) return x.id < y.id;
        return x.pos < y.pos;
    });

    // assignGroup (kmermatcher.cpp:450-559), statement for statement.
    size_t writePos = 0;
    {
        or_kmer_rec *hp = a.data();
        uint64_t prevHash = hp[0].kmer;
        uint64_t repSeqId = hp[0].id;
        if (nt) {
            bool isReverse = (bitCheck63(hp[0].kmer) == false);
            repSeqId = isReverse ? bitClear63(repSeqId) : bitSet63(repSeqId);
            prevHash = bitSet63(prevHash);
        }
        size_t prevHashStart = 0, prevSetSize = 0;
        int32_t queryLen = hp[0].seq_len;
        bool repIsReverse = false;
        int32_t repPos = hp[0].pos;
        for (size_t e = 0; e < totalKmersPerSplit + 1; e++) {
            uint64_t currKmer = hp[e].kmer;
            if (nt) currKmer = bitSet63(currKmer);
            if (prevHash != currKmer) {
                for (size_t i = prevHashStart; i < e; i++) {
                    uint64_t kmer = hp[i].kmer;
                    if (nt) kmer = bitSet63(hp[i].kmer);
                    uint64_t rId = (kmer != U64MAX) ? ((prevSetSize == 1) ? U64MAX : repSeqId) : U64MAX;
                    if (rId != U64MAX) {
                        int diagonal = repPos - hp[i].pos;
                        if (nt) {
                            bool targetIsReverse = (bitCheck63(hp[i].kmer) == false);
                            bool queryNeedsToBeRev = false;
                            int32_t queryPos = 0, targetPos = 0;
                            if (repIsReverse == true && targetIsReverse == false) {
                                queryPos = repPos; targetPos = hp[i].pos; queryNeedsToBeRev = true;
                            } else if (repIsReverse == true && targetIsReverse == true) {
                                queryPos = tshort((queryLen - 1) - repPos);
                                targetPos = tshort((hp[i].seq_len - 1) - hp[i].pos);
                                queryNeedsToBeRev = false;
                            } else if (repIsReverse == false && targetIsReverse == true) {
                                queryPos = tshort((queryLen - 1) - repPos);
                                targetPos = tshort((hp[i].seq_len - 1) - hp[i].pos);
                                queryNeedsToBeRev = true;
                            } else {
                                queryPos = repPos; targetPos = hp[i].pos; queryNeedsToBeRev = false;
                            }
                            diagonal = queryPos - targetPos;
                            rId = queryNeedsToBeRev ? bitClear63(rId) : bitSet63(rId);
                        }
                        bool canBeExtended = diagonal < 0 || (diagonal > (queryLen - hp[i].seq_len));
                        bool covered = canBeCovered(p->cov_thr, p->cov_mode, (float) queryLen, (float) hp[i].seq_len);
                        if ((p->include_only_extendable == 0 && covered) || (canBeExtended && p->include_only_extendable != 0)) {
                            int32_t sl = hp[i].seq_len; uint32_t id = hp[i].id;
                            hp[writePos].kmer = rId;
                            hp[writePos].pos = tshort(diagonal);
                            hp[writePos].seq_len = sl;
                            hp[writePos].id = id;
                            writePos++;
                        }
                    }
                    hp[i].kmer = (i != writePos - 1) ? U64MAX : hp[i].kmer;
                }
                prevSetSize = 0;
                prevHashStart = e;
                repSeqId = hp[e].id;
                if (nt) {
                    repIsReverse = (bitCheck63(hp[e].kmer) == 0);
                    repSeqId = repIsReverse ? repSeqId : bitSet63(repSeqId);
                }
                queryLen = hp[e].seq_len;
                repPos = hp[e].pos;
            }
            if (hp[e].kmer == U64MAX) break;
            prevSetSize++;
            prevHash = hp[e].kmer;
            if (nt) prevHash = bitSet63(prevHash);
        }
    }

    // sort #2 (kmermatcher.cpp:427-431; kmermatcher.h:98-130)
    std::sort(a.begin(), a.begin() + writePos, [nt](const or_kmer_rec &x, const or_kmer_rec &y) {
        uint64_t kx = nt ? bitSet63(x.kmer) : x.kmer, ky = nt ? bitSet63(y.kmer) : y.kmer;
        if (kx != ky) return kx < ky;
        if (x.id != y.id) return x.id < y.id;
        return x.pos < y.pos;
    });

    // writeKmerMatcherResult with threads = 1 (kmermatcher.cpp:809-924).  Only hit lines are
    // recorded: a block "rep\t0\t0\n" without hits is byte-identical to the default entry (:705-724).
    std::vector<or_hit> hits;
    {
        const or_kmer_rec *hp = a.data();
        const size_t total = totalKmersPerSplit;
        uint64_t lastTargetId = U64MAX;
        uint64_t repSeqId = U64MAX;
        for (size_t kmerPos = 0; kmerPos < total && hp[kmerPos].kmer != U64MAX; kmerPos++) {
            uint64_t currKmer = hp[kmerPos].kmer;
            int reverMask = 0;
            if (nt) { reverMask = bitCheck63(currKmer) == false; currKmer = bitClear63(currKmer); }
            if (repSeqId != currKmer) {
                lastTargetId = U64MAX;
                repSeqId = currKmer;
            }
            unsigned int targetId = hp[kmerPos].id;
            int32_t diagonal = hp[kmerPos].pos;
            size_t kmerOffset = 0;
            int32_t prevDiagonal = diagonal;
            size_t maxDiagonal = 0, diagonalCnt = 0, topScore = 0;
            int bestReverMask = reverMask;
            while (lastTargetId != targetId && kmerPos + kmerOffset < total && hp[kmerPos + kmerOffset].id == targetId) {
                if (prevDiagonal == hp[kmerPos + kmerOffset].pos) diagonalCnt++; else diagonalCnt = 1;
                if (diagonalCnt >= maxDiagonal) {
                    diagonal = hp[kmerPos + kmerOffset].pos;
                    maxDiagonal = diagonalCnt;
                    if (nt) bestReverMask = bitCheck63(hp[kmerPos + kmerOffset].kmer) == false;
                }
                prevDiagonal = hp[kmerPos + kmerOffset].pos;
                kmerOffset++;
                topScore++;
            }
            if (targetId != repSeqId && lastTargetId != targetId) {
                ;
            } else {
                lastTargetId = targetId;
                continue;
            }
            or_hit h;
            h.rep = (uint32_t) repSeqId;
            h.target = targetId;
            h.score = bestReverMask ? -(int32_t) topScore : (int32_t) topScore;
            h.diag = (int32_t) (short) (unsigned short) diagonal;   // hit_t::diagonal is u16, printed as short
            hits.push_back(h);
            lastTargetId = targetId;
        }
    }
    *n_out = hits.size();
    *out = (or_hit *) malloc(sizeof(or_hit) * (hits.size() + 1));
    if (!hits.empty()) memcpy(*out, hits.data(), sizeof(or_hit) * hits.size());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// E-values: EvalueComputation (mm/alignment/EvalueComputation.h:18-40) over ALP
// (lib/mmseqs/lib/alp/sls_alignment_evaluer.hpp:150-167, sls_pvalues.cpp:342-525,
//  sls_basic.hpp:195-198 normal_probability(x) = 0.5*erfc(-sqrt(0.5)*x)).
// ---------------------------------------------------------------------------------------------
namespace {
struct Alp {
    double lambda, K, a_I, b_I, a_J, b_J, alpha_I, beta_I, alpha_J, beta_J, sigma, tau;
    double vi_y_thr, vj_y_thr, c_y_thr, logK;
};
Alp makeAlp(bool nt) {
    const double *t = nt ? OR_NT_ALP : OR_AA_ALP;
    Alp p;
    p.lambda = t[0]; p.K = t[1]; p.a_I = t[2]; p.b_I = t[3]; p.a_J = t[4]; p.b_J = t[5];
    p.alpha_I = t[6]; p.beta_I = t[7]; p.alpha_J = t[8]; p.beta_J = t[9]; p.sigma = t[10]; p.tau = t[11];
    // pvalues::compute_tmp_values (sls_pvalues.cpp:342-364), nat_cut_off_in_max = 2.0
    p.vi_y_thr = std::max(2.0 * p.alpha_I / p.lambda, 0.0);
    p.vj_y_thr = std::max(2.0 * p.alpha_J / p.lambda, 0.0);
    p.c_y_thr = std::max(2.0 * p.sigma / p.lambda, 0.0);
    p.logK = log(p.K);
    return p;
}
const double ALP_PI = 3.1415926535897932384626433832795;
inline double normalProbability(double x) { return 0.5 * erfc(-sqrt(0.5) * x); }

// pvalues::get_appr_tail_prob_with_cov_without_errors with compute_only_area (sls_pvalues.cpp:366-525),
// called as area(score, seqlen1 = qLen, seqlen2 = dbRes) => m_ = seqlen2, n_ = seqlen1
// (sls_alignment_evaluer.cpp:989-1025).
double alpArea(const Alp &p, double y, double seqlen1, double seqlen2) {
    const double const_val = 1 / sqrt(2.0 * ALP_PI);
    double m_ = seqlen2, n_ = seqlen1;
    double m_li_y = m_ - (p.a_I * y + p.b_I);
    double vi_y = std::max(p.vi_y_thr, p.alpha_I * y + p.beta_I);
    double sqrt_vi_y = sqrt(vi_y);
    double m_F = (sqrt_vi_y == 0.0) ? 1e100 : m_li_y / sqrt_vi_y;
    double P_m_F = normalProbability(m_F);
    double E_m_F = -const_val * exp(-0.5 * m_F * m_F);
    double p1 = m_li_y * P_m_F - sqrt_vi_y * E_m_F;

    double n_lj_y = n_ - (p.a_J * y + p.b_J);
    double vj_y = std::max(p.vj_y_thr, p.alpha_J * y + p.beta_J);
    double sqrt_vj_y = sqrt(vj_y);
    double n_F = (sqrt_vj_y == 0.0) ? 1e100 : n_lj_y / sqrt_vj_y;
    double P_n_F = normalProbability(n_F);
    double E_n_F = -const_val * exp(-0.5 * n_F * n_F);
    double p2 = n_lj_y * P_n_F - sqrt_vj_y * E_n_F;

    double c_y = std::max(p.c_y_thr, p.sigma * y + p.tau);
    double area = p1 * p2 + c_y * (P_m_F * P_n_F);
    return area;
}
struct Evaluer {
    Alp p; double dbRes;
    Evaluer(bool nt, double dbRes) : p(makeAlp(nt)), dbRes(dbRes) {}
    double evalue(double score, double qLen) const {           // EvalueComputation.h:36-40
        double epa = p.K * exp(-p.lambda * score);
        double a = alpArea(p, score, qLen, dbRes);
        return epa * a;
    }
    double bitScore(double score) const { return (p.lambda * score - p.logK) / log(2.0); }   // :18-20
    double rawFromBits(double bits) const { return (p.logK + bits * std::log(2.0)) / p.lambda; }  // :22-24
};

// SubstitutionMatrix::createAsciiSubMat (mm/commons/SubstitutionMatrix.h:56-73): 123 x 123 by raw ASCII.
struct AsciiMat {
    signed char m[123][123];
    AsciiMat(bool nt) {
        const unsigned char *a2n = nt ? OR_NT_AA2NUM : OR_AA_AA2NUM;
        const signed char *sm = nt ? OR_NT_SUBMAT : OR_AA_SUBMAT;
        const int A = nt ? 5 : 21;
        for (int i = 0; i < 123; i++) for (int j = 0; j < 123; j++) m[i][j] = sm[a2n[i] * A + a2n[j]];
    }
};

struct LocalAln {   // DistanceCalculator::LocalAlignment (mm/alignment/DistanceCalculator.h:42-55)
    int startPos = -1, endPos = -1;
    unsigned int score = 0, diagonalLen = 0, distToDiagonal = 0;
    int diagonal = 0;
};

// computeGlobalSubstitutionStartEndDistance (DistanceCalculator.h:204-220)
void globalScore(const char *s1, const char *s2, unsigned int length, const AsciiMat &M, LocalAln &res) {
    unsigned int first = (s1[0] == '*' || s2[0] == '*') ? 1 : 0;
    unsigned int last = length - 1;
    if (last > 0 && (s1[length - 1] == '*' || s2[length - 1] == '*')) last--;
    int64_t score = 0;
    for (unsigned int pos = first; pos <= last; pos++) score += M.m[(int) s1[pos]][(int) s2[pos]];
    score = std::max(score, (int64_t) 0);
    res.startPos = (int) first; res.endPos = (int) last; res.score = (unsigned int) score;
}

// ungappedAlignmentByDiagonal, mode 3 only (DistanceCalculator.h:115-175)
LocalAln alignByDiagonal(const char *q, unsigned int qLen, const char *t, unsigned int tLen, int diagonal, const AsciiMat &M) {
    unsigned int dist = (unsigned int) abs(diagonal);
    LocalAln res;
    res.distToDiagonal = dist;
    res.diagonal = diagonal;
    if (diagonal >= 0 && dist < qLen) {
        unsigned int minSeqLen = std::min(tLen, qLen - dist);
        res.diagonalLen = minSeqLen;
        globalScore(q + dist, t, minSeqLen, M, res);
    } else if (diagonal < 0 && dist < tLen) {
        unsigned int minSeqLen = std::min(tLen - dist, qLen);
        res.diagonalLen = minSeqLen;
        globalScore(q, t + dist, minSeqLen, M, res);
    }
    return res;
}

// computeUngappedAlignment (DistanceCalculator.h:94-113): all 65536-wrapped candidates of a u16 diagonal
LocalAln computeUngapped(const char *q, unsigned int qLen, const char *t, unsigned int tLen, unsigned short diagonal, const AsciiMat &M) {
    LocalAln max;
    for (unsigned int d = 1; d <= 1 + tLen / 32768; d++) {
        int real = (int) (-d * 65536 + diagonal);
        LocalAln tmp = alignByDiagonal(q, qLen, t, tLen, real, M);
        if (tmp.score > max.score) max = tmp;
    }
    for (unsigned int d = 0; d <= qLen / 65536; d++) {
        int real = (int) (d * 65536 + diagonal);
        LocalAln tmp = alignByDiagonal(q, qLen, t, tLen, real, M);
        if (tmp.score > max.score) max = tmp;
    }
    return max;
}

// SmithWaterman::computeCov (mm/alignment/StripedSmithWaterman.cpp:1055-1057), unsigned arithmetic
float computeCov(unsigned int s, unsigned int e, unsigned int len) {
    return (std::min(len, std::max(s, e)) - std::min(s, e) + 1) / (float) len;
}

// plain decimal, as Itoa::i32toa_sse2 / u32toa_sse2 produce
char *putInt(char *b, long long v) { return b + sprintf(b, "%lld", v); }

// Util::fastSeqIdToBuffer (mm/commons/Util.cpp:278-307) without the trailing separator
char *putSeqId(char *b, float seqId) {
    // the reference writes "1.000\0" but returns the pointer to the NUL, whose predecessor the caller
    // overwrites with the separator (Matcher.cpp:330-331) => "1.00" on disk
    if (seqId == 1.0) { memcpy(b, "1.00", 4); return b + 4; }
    *b++ = '0'; *b++ = '.';
    if (seqId < 0.10) *b++ = '0';
    if (seqId < 0.01) *b++ = '0';
    return putInt(b, (int) (seqId * 1000));
}

// nucleotide reverse complement by letter (rescorediagonal.cpp:173-179: num2aa[reverse[aa2num[c]]])
inline char ntRevLetter(char c) { return (char) OR_NT_NUM2AA[OR_NT_REVERSE[OR_NT_AA2NUM[(unsigned char) c]]]; }

}  // namespace

extern "C" double or_evalue(int nt, double db_residues, double score, double q_len) { return Evaluer(nt != 0, db_residues).evalue(score, q_len); }
extern "C" double or_bitscore(int nt, double score) { return Evaluer(nt != 0, 1).bitScore(score); }
extern "C" double or_raw_from_bits(int nt, double bits) { return Evaluer(nt != 0, 1).rawFromBits(bits); }

extern "C" int or_format_hit(char *buf, uint32_t target, int32_t score, int32_t diag) {
    char *b = putInt(buf, target); *b++ = '\t';
    b = putInt(b, score); *b++ = '\t';
    b = putInt(b, (short) diag); *b++ = '\n'; *b = '\0';
    return (int) (b - buf);
}

extern "C" int or_format_aln(char *buf, const or_aln *a) {
    char *b = putInt(buf, a->target); *b++ = '\t';
    b = putInt(b, a->bits); *b++ = '\t';
    b = putSeqId(b, a->seq_id); *b++ = '\t';
    b += sprintf(b, "%.3E", a->evalue); *b++ = '\t';
    b = putInt(b, a->q_start); *b++ = '\t';
    b = putInt(b, a->q_end); *b++ = '\t';
    b = putInt(b, a->q_len); *b++ = '\t';
    b = putInt(b, a->db_start); *b++ = '\t';
    b = putInt(b, a->db_end); *b++ = '\t';
    b = putInt(b, a->db_len); *b++ = '\n'; *b = '\0';
    return (int) (b - buf);
}

// rescorediagonal for one prefilter line (rescorediagonal.cpp:193-330), query DB == target DB.
extern "C" int or_rescore(const or_seqdb *db, const or_hit *hits, uint64_t n_hits, const or_rs_params *p,
                          or_aln **out, uint64_t *n_out) {
    if (p->rescore_mode != 3) return -1;
    const bool nt = db->dbtype == 1;
    AsciiMat M(nt);
    double dbRes = 0;                     // getAminoAcidDBSize = sum(len) - 2N (mm/commons/DBReader.cpp:537-546)
    for (uint64_t i = 0; i < db->n; i++) dbRes += (double) db->lens[i] - 2;
    Evaluer ev(nt, dbRes);
    std::vector<or_aln> res;
    std::string qRev;
    uint64_t hp = 0;
    // hits are grouped by rep in ascending rep order; keys ascending.
    for (uint64_t qi = 0; qi < db->n; qi++) {
        Seq q = getSeq(db, qi);
        if (nt) {                         // full reverse complement of the query (:173-179)
            qRev.resize(q.len);
            for (int pos = q.len - 1; pos > -1; pos--) qRev[(q.len - 1) - pos] = ntRevLetter(q.s[pos]);
        }
        while (hp < n_hits && hits[hp].rep < q.key) hp++;
        uint64_t he = hp;
        while (he < n_hits && hits[he].rep == q.key) he++;
        for (uint64_t li = 0; li < 1 + (he - hp); li++) {
            uint32_t tKey; int prefScore; unsigned short diag16;
            if (li == 0) { tKey = q.key; prefScore = 0; diag16 = 0; }           // the "rep\t0\t0" line
            else { tKey = hits[hp + li - 1].target; prefScore = hits[hp + li - 1].score; diag16 = (unsigned short) (short) hits[hp + li - 1].diag; }
            const char *qAlign = q.s;
            bool isReverse = false;
            if (nt && prefScore < 0) { qAlign = qRev.c_str(); isReverse = true; }
            uint64_t ti = findId(db, tKey);
            if (ti == U64MAX) return -2;
            const bool isIdentity = (qi == ti);       // sameQTDB (:205)
            Seq t = getSeq(db, ti);
            int dbLen = t.len, qLen = q.len;
            if (!canBeCovered(p->cov_thr, p->cov_mode, (float) qLen, (float) dbLen)) continue;
            LocalAln al = computeUngapped(qAlign, (unsigned) qLen, t.s, (unsigned) (float) dbLen, diag16, M);
            unsigned int distToDiag = al.distToDiagonal;
            int distance = (int) al.score;
            int diagonal = al.diagonal;
            double seqId = 0;
            double evalue = ev.evalue(distance, qLen);
            int bitScore = static_cast<int>(ev.bitScore(distance) + 0.5);
            int alnLen = (al.endPos - al.startPos) + 1;
            int qS, qE, dS, dE;
            if (diagonal >= 0) { qS = al.startPos + distToDiag; qE = al.endPos + distToDiag; dS = al.startPos; dE = al.endPos; }
            else { qS = al.startPos; qE = al.endPos; dS = al.startPos + distToDiag; dE = al.endPos + distToDiag; }
            if (evalue <= p->eval_thr || isIdentity) {
                int idCnt = 0;
                for (int i = qS; i <= qE; i++) {
                    // hazard #12: for a zero-score hit qS = -1 and the reference reads the byte
                    // before both sequences; for the identity hit both pointers are the same byte => equal.
                    char ql = (i < 0) ? 0 : (qAlign[i] & (unsigned char) ~0x20);
                    int tp = dS + (i - qS);
                    char tl = (tp < 0) ? 0 : (t.s[tp] & (unsigned char) ~0x20);
                    idCnt += (ql == tl) ? 1 : 0;
                }
                // Util::computeSeqId (Util.cpp:588-598)
                float s;
                if (p->seq_id_mode == 1) s = (float) idCnt / (float) std::min(qLen, dbLen);
                else if (p->seq_id_mode == 2) s = (float) idCnt / (float) std::max(qLen, dbLen);
                else s = (float) idCnt / (float) alnLen;
                seqId = s;
            }
            float queryCov = computeCov((unsigned) qS, (unsigned) qE, (unsigned) qLen);
            float targetCov = computeCov((unsigned) dS, (unsigned) dE, (unsigned) dbLen);
            if (isReverse) { qS = qLen - qS - 1; qE = qLen - qE - 1; }
            bool hasCov = hasCoverage(p->cov_thr, p->cov_mode, queryCov, targetCov);
            bool hasSeqId = seqId >= (p->seq_id_thr - std::numeric_limits<float>::epsilon());
            bool hasEvalue = (evalue <= p->eval_thr);
            bool hasAlnLen = (alnLen >= p->aln_len_thr);
            if (isIdentity || (hasAlnLen && hasCov && hasSeqId && hasEvalue)) {
                or_aln a;
                a.query = q.key; a.target = tKey; a.bits = bitScore; a.seq_id = (float) seqId; a.evalue = evalue;
                a.q_start = qS; a.q_end = qE; a.q_len = qLen; a.db_start = dS; a.db_end = dE; a.db_len = dbLen;
                res.push_back(a);
            }
        }
        hp = he;
    }
    *n_out = res.size();
    *out = (or_aln *) malloc(sizeof(or_aln) * (res.size() + 1));
    if (!res.empty()) memcpy(*out, res.data(), sizeof(or_aln) * res.size());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// assembleresults / nuclassembleresults
// ---------------------------------------------------------------------------------------------
namespace {
struct Res {   // Matcher::result_t subset
    uint32_t dbKey; int score; float seqId; unsigned int alnLength;
    int qStartPos, qEndPos; unsigned int qLen; int dbStartPos, dbEndPos; unsigned int dbLen;
};
struct CmpAa {   // CompareResultByScore (assembleresult.cpp:19-36)
    bool operator()(const Res &r1, const Res &r2) const {
        if (r1.score < r2.score) return true;
        if (r2.score < r1.score) return false;
        if (r1.alnLength < r2.alnLength) return true;
        if (r2.alnLength < r1.alnLength) return false;
        if (r1.dbKey > r2.dbKey) return true;
        if (r2.dbKey > r1.dbKey) return false;
        return false;
    }
};
struct CmpNt {   // CompareNuclResultByScore (nuclassembleresult.cpp:36-70)
    bool operator()(const Res &r1, const Res &r2) const {
        unsigned int mm_count1 = (1 - r1.seqId) * r1.alnLength + 0.5;
        unsigned int mm_count2 = (1 - r2.seqId) * r2.alnLength + 0.5;
        unsigned int alpha1 = mm_count1 + 1;
        unsigned int alpha2 = mm_count2 + 1;
        unsigned int beta1 = r1.alnLength - mm_count1 + 1;
        unsigned int beta2 = r2.alnLength - mm_count2 + 1;
        double log_c = (std::lgamma(beta1 + beta2) + std::lgamma(alpha1 + beta1)) - (std::lgamma(alpha1 + beta1 + beta2) + std::lgamma(beta1));
        double log_r = 0.0;
        double p = 0.0;
        for (size_t idx = 0; idx < alpha2; idx++) {
            p += exp(log_r + log_c);
            log_r = log(alpha1 + idx) + log(beta2 + idx) - (log(idx + 1) + log(idx + alpha1 + beta1 + beta2)) + log_r;
        }
        if (p < 0.45) return true;
        if (p > 0.55) return false;
        if (r1.dbLen - r1.alnLength < r2.dbLen - r2.alnLength) return true;
        if (r1.dbLen - r1.alnLength > r2.dbLen - r2.alnLength) return false;
        return true;
    }
};

// getRevFragment / getNuclRevFragment (assembleresult.cpp:59-68, nuclassembleresult.cpp:93-102)
std::string revFragment(const char *frag, size_t len) {
    std::string r(len, ' ');
    for (int pos = (int) len - 1; pos > -1; pos--) {
        char rv = ntRevLetter(frag[pos]);
        r[(len - 1) - pos] = (rv == 'X') ? 'N' : rv;
    }
    return r;
}

// updateAlignment / updateNuclAlignment (assembleresult.cpp:70-108)
void updateAlignment(Res &r, const LocalAln &al, const char *q, size_t qLen, const char *t, size_t tLen) {
    int diag = al.diagonal;
    int dist = std::max(abs(diag), 0);
    int qS, qE, dS, dE;
    if (diag >= 0) { qS = al.startPos + dist; qE = al.endPos + dist; dS = al.startPos; dE = al.endPos; }
    else { qS = al.startPos; qE = al.endPos; dS = al.startPos + dist; dE = al.endPos + dist; }
    int idCnt = 0;
    for (int i = qS; i < qE; i++) idCnt += (q[i] == t[dS + (i - qS)]) ? 1 : 0;
    float seqId = static_cast<float>(idCnt) / (static_cast<float>(qE) - static_cast<float>(qS));
    r.seqId = seqId;
    r.qLen = (unsigned int) qLen;
    r.dbLen = (unsigned int) tLen;
    r.alnLength = al.diagonalLen;
    float scorePerCol = static_cast<float>(al.score) / static_cast<float>(r.alnLength + 0.5);
    r.score = static_cast<int>(scorePerCol * 100);
    r.qStartPos = qS; r.qEndPos = qE; r.dbStartPos = dS; r.dbEndPos = dE;
}

template <class Cmp>
bool selectFragment(std::priority_queue<Res, std::vector<Res>, Cmp> &qu, unsigned int queryKey, Res &out) {
    // selectFragmentToExtend (assembleresult.cpp:40-57)
    while (!qu.empty()) {
        Res res = qu.top();
        qu.pop();
        const bool notRightStartAndLeftStart = !(res.dbStartPos == 0 && res.qStartPos == 0);
        const bool rightStart = res.dbStartPos == 0 && (res.dbEndPos != static_cast<int>(res.dbLen) - 1);
        const bool leftStart = res.qStartPos == 0 && (res.qEndPos != static_cast<int>(res.qLen) - 1);
        const bool isNotIdentity = (res.dbKey != queryKey);
        if ((rightStart || leftStart) && notRightStartAndLeftStart && isNotIdentity) { out = res; return true; }
    }
    return false;
}

template <class Cmp>
bool extendOne(const or_seqdb *db, uint64_t id, const or_aln *alns, uint64_t nAl, const or_ex_params *p,
               bool nt, const AsciiMat &M, const Evaluer &ev, std::vector<char> &useReverse,
               std::vector<unsigned char> &wasExtended, std::string &query) {
    Seq qs = getSeq(db, id);
    unsigned int queryKey = qs.key;
    const char *querySeq = qs.s;
    unsigned int querySeqLen = (unsigned int) qs.len;
    query.assign(querySeq, querySeqLen);
    bool queryCouldBeExtended = false;
    std::priority_queue<Res, std::vector<Res>, Cmp> alnQueue;
    for (uint64_t i = 0; i < nAl; i++) {
        // Matcher::parseAlignmentRecord after the text round trip (Matcher.cpp:248-320, :323-370)
        char buf[64];
        *putSeqId(buf, alns[i].seq_id) = '\0';
        Res r;
        r.dbKey = alns[i].target;
        r.score = alns[i].bits;
        r.seqId = (float) strtod(buf, NULL);
        r.qStartPos = alns[i].q_start; r.qEndPos = alns[i].q_end; r.qLen = (unsigned) alns[i].q_len;
        r.dbStartPos = alns[i].db_start; r.dbEndPos = alns[i].db_end; r.dbLen = (unsigned) alns[i].db_len;
        int adjQ = (r.qStartPos == -1) ? 0 : r.qStartPos;
        int adjD = (r.dbStartPos == -1) ? 0 : r.dbStartPos;
        r.alnLength = (unsigned int) (std::max(abs(r.qEndPos - adjQ), abs(r.dbEndPos - adjD)) + 1);
        // fill queue (assembleresult.cpp:159-188 / nuclassembleresult.cpp:197-226)
        int rawScore = static_cast<int>(ev.rawFromBits(r.score) + 0.5);
        float scorePerCol = static_cast<float>(rawScore) / static_cast<float>(r.alnLength + 0.5);
        if (!nt) {
            float alnLen = static_cast<float>(r.alnLength);
            float ids = static_cast<float>(r.seqId) * alnLen;
            r.seqId = ids / (alnLen + 0.5);
        }
        r.score = static_cast<int>(scorePerCol * 100);
        if (nt) {
            uint64_t tid = findId(db, r.dbKey);
            if (r.qStartPos > r.qEndPos) {
                useReverse[tid] = true;
                std::swap(r.qStartPos, r.qEndPos);
                unsigned int dbStartPos = r.dbStartPos;
                r.dbStartPos = r.dbLen - r.dbEndPos - 1;
                r.dbEndPos = r.dbLen - dbStartPos - 1;
            } else {
                useReverse[tid] = false;
            }
        }
        alnQueue.push(r);
    }
    std::vector<Res> tmpAlignments;
    while (!alnQueue.empty()) {
        unsigned int leftQueryOffset = 0, rightQueryOffset = 0;
        tmpAlignments.clear();
        Res best;
        bool broke = false;
        while (selectFragment(alnQueue, queryKey, best)) {
            uint64_t targetId = findId(db, best.dbKey);
            Seq ts = getSeq(db, targetId);
            const char *targetSeq = ts.s;
            unsigned int targetSeqLen = (unsigned int) ts.len;
            if (best.dbStartPos == 0) {
                if ((targetSeqLen - (best.dbEndPos + 1)) <= rightQueryOffset) continue;
            } else if (best.qStartPos == 0) {
                if (best.dbStartPos <= static_cast<int>(leftQueryOffset)) continue;
            }
            unsigned int dbStartPos = best.dbStartPos, dbEndPos = best.dbEndPos;
            unsigned int qStartPos = best.qStartPos, qEndPos = best.qEndPos;
            if (dbStartPos == 0 && qEndPos == (querySeqLen - 1)) {          // right extension
                if (rightQueryOffset > 0) { tmpAlignments.push_back(best); continue; }
                unsigned int fragLen = targetSeqLen - (dbEndPos + 1);
                if (nt && query.size() + fragLen >= (size_t) p->max_seq_len) { broke = true; break; }   // nucl only (:271-275)
                std::string fragment;
                if (useReverse[targetId]) fragment = revFragment(targetSeq, fragLen);
                else fragment = std::string(targetSeq + dbEndPos + 1, fragLen);
                query += fragment;
                rightQueryOffset += fragLen;
                wasExtended[targetId] |= 0x80;
            } else if (qStartPos == 0 && dbEndPos == (targetSeqLen - 1)) {   // left extension
                if (leftQueryOffset > 0) { tmpAlignments.push_back(best); continue; }
                unsigned int fragLen = dbStartPos;
                if (query.size() + fragLen >= (size_t) p->max_seq_len) { broke = true; break; }
                std::string fragment;
                if (useReverse[targetId]) fragment = revFragment(targetSeq + (targetSeqLen - dbStartPos), fragLen);
                else fragment = std::string(targetSeq, fragLen);
                query = fragment + query;
                leftQueryOffset += fragLen;
                wasExtended[targetId] |= 0x80;
            }
        }
        (void) broke;
        if (leftQueryOffset > 0 || rightQueryOffset > 0) queryCouldBeExtended = true;
        if (!alnQueue.empty()) break;
        querySeqLen = (unsigned int) query.length();
        querySeq = query.c_str();
        for (size_t ai = 0; ai < tmpAlignments.size(); ai++) {
            uint64_t tId = findId(db, tmpAlignments[ai].dbKey);
            Seq ts = getSeq(db, tId);
            unsigned int tSeqLen = (unsigned int) ts.len;
            const char *tSeq = ts.s;
            std::string rv;
            if (useReverse[tId]) { rv = revFragment(tSeq, tSeqLen); tSeq = rv.c_str(); }
            int qStartPos = tmpAlignments[ai].qStartPos;
            int dbStartPos = tmpAlignments[ai].dbStartPos;
            int diag = (qStartPos + leftQueryOffset) - dbStartPos;
            LocalAln al = alignByDiagonal(querySeq, querySeqLen, tSeq, tSeqLen, diag, M);
            updateAlignment(tmpAlignments[ai], al, querySeq, querySeqLen, tSeq, tSeqLen);
            if (tmpAlignments[ai].seqId >= p->seq_id_thr) alnQueue.push(tmpAlignments[ai]);
        }
    }
    return queryCouldBeExtended;
}
}  // namespace

extern "C" int or_extend(const or_seqdb *db, const or_aln *alns, uint64_t n_alns, const or_ex_params *p,
                         char **out_data, uint64_t **out_offsets, uint32_t **out_lens, uint32_t **out_keys,
                         uint8_t **extended, uint64_t *out_n, uint64_t *out_bytes) {
    if (p->rescore_mode != 3) return -1;
    const bool nt = db->dbtype == 1;
    AsciiMat M(nt);
    double dbRes = 0;
    for (uint64_t i = 0; i < db->n; i++) dbRes += (double) db->lens[i] - 2;
    Evaluer ev(nt, dbRes);
    std::vector<char> useReverse(db->n, 0);
    std::vector<unsigned char> wasExtended(db->n, 0);
    std::vector<std::string> contigs(db->n);
    uint64_t ap = 0;
    std::string query;
    for (uint64_t id = 0; id < db->n; id++) {
        uint32_t key = db->keys[id];
        while (ap < n_alns && alns[ap].query < key) ap++;
        uint64_t ae = ap;
        while (ae < n_alns && alns[ae].query == key) ae++;
        bool extd = nt ? extendOne<CmpNt>(db, id, alns + ap, ae - ap, p, nt, M, ev, useReverse, wasExtended, query)
                       : extendOne<CmpAa>(db, id, alns + ap, ae - ap, p, nt, M, ev, useReverse, wasExtended, query);
        ap = ae;
        if (extd) { wasExtended[id] |= 0x20; contigs[id] = query; }
    }
    std::string all;
    uint64_t *offs = (uint64_t *) malloc(sizeof(uint64_t) * (db->n + 1));
    uint32_t *lens = (uint32_t *) malloc(sizeof(uint32_t) * (db->n + 1));
    uint32_t *keys = (uint32_t *) malloc(sizeof(uint32_t) * (db->n + 1));
    uint8_t *ext = (uint8_t *) malloc(db->n + 1);
    uint64_t w = 0;
    for (uint64_t id = 0; id < db->n; id++) {
        bool isContig = wasExtended[id] & 0x20;
        bool wasNotExtended = !(wasExtended[id] & 0x80);
        if (isContig) {                       // assembleresult.cpp:316-320
            offs[w] = all.size();
            all.append(contigs[id]); all.push_back('\n'); all.push_back('\0');
            lens[w] = (uint32_t) contigs[id].size() + 2;
        } else if (p->keep_target || wasNotExtended) {   // :326-342
            offs[w] = all.size();
            all.append(db->data + db->offsets[id], db->lens[id] - 1); all.push_back('\0');
            lens[w] = db->lens[id];
        } else {
            continue;
        }
        keys[w] = db->keys[id];
        ext[w] = isContig ? 1 : 0;
        w++;
    }
    *out_data = (char *) malloc(all.size() + 1);
    memcpy(*out_data, all.data(), all.size());
    *out_offsets = offs; *out_lens = lens; *out_keys = keys; *extended = ext; *out_n = w; *out_bytes = all.size();
    return 0;
}

extern "C" void or_free(void *p) { free(p); }
