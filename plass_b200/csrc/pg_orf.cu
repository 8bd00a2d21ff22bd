// pg_orf.cu -- six-frame ORF extraction and translation of the reads: the step that turns
// nucl_reads into the amino-acid fragments the first assemble iteration starts from (data/assemble.sh:41-77).
//
//   orf_run  replaces extractorfs                       lib/mmseqs/src/util/extractorfs.cpp:20-159 (Orf::findForward,
//                                                       lib/mmseqs/src/commons/Orf.cpp:192-330)
//            and, with translate != 0, the translatenucs --add-orf-stop 1 that follows it
//                                                       lib/mmseqs/src/util/translatenucs.cpp:14-128 (TranslateNucl.h:333-503)
//
// One thread per read walks the codons of both strands in the reference's order (position-major, frame = position % 3)
// with the three per-frame state machines in registers; a counting pass sizes the output, two scans place every read's
// fragments, an emitting pass writes the fragments, their keys (0..n-1 in (read, emission) order = the renumbered DB of
// DBWriter::createRenumberedDB) and the ORF header fields.  Byte / integer work, no tensor cores.
#include "pg_internal.cuh"
#include "pg_scan.cuh"

#include <algorithm>
#include <cstring>
#include <string>

namespace pg {

struct OrfConst {
    int minLength, maxLength, maxGaps, contigStartMode, contigEndMode, startMode;
    unsigned forwardFrames, reverseFrames;
    int allStarts, translate;
};

__constant__ char c_orf_rc[256];             // Orf::iupacReverseComplementTable (Orf.cpp:48-52)
__constant__ unsigned char c_orf_b2i[256];   // TranslateNucl::sm_BaseToIdx
// Codon tests of Orf::findForward as integer compares: every character is mapped ONCE to a 3-bit class -- 0..3 = A C G T
// (after the upper-casing of :226-229), 4 = any other IUPAC letter, 5 = 'N' or a character whose complement is '.'
// (isGapOrN, :181-185), 7 = the CHAR_MAX padding -- by one table for the forward strand and one for the reverse strand
// (class of the complemented character, '.' -> 'N').  A codon is c0 * 64 + c1 * 8 + c2.
__constant__ unsigned char c_orf_clsF[256];
__constant__ unsigned char c_orf_clsR[256];
constexpr unsigned ORF_ATG = 0 * 64 + 3 * 8 + 2, ORF_TTG = 3 * 64 + 3 * 8 + 2, ORF_CTG = 1 * 64 + 3 * 8 + 2;
constexpr unsigned ORF_TAA = 3 * 64 + 0 * 8 + 0, ORF_TAG = 3 * 64 + 0 * 8 + 2, ORF_TGA = 3 * 64 + 2 * 8 + 0;

constexpr char ORF_PAD = 127;                // CHAR_MAX: the padding behind the sequence (Orf.cpp:158-161)

// character p of the strand: forward = the read ('u' -> 't' only: the 'U' branch of Orf::setSequence is overwritten,
// Orf.cpp:144-147), reverse = complement of the read backwards with '.' -> 'N' (:149-155)
// The two 256-byte tables are read with a different index in every lane: the kernels copy them from constant to shared
// memory (the constant cache serialises divergent addresses).
struct OrfLut { const char *rc; const unsigned char *b2i; const unsigned char *clsF, *clsR; };

__device__ __forceinline__ unsigned orf_class(const char *__restrict__ seq, unsigned L, unsigned p, bool reverse, const OrfLut &lut) {
    if (p >= L) return 7u;
    return reverse ? lut.clsR[(unsigned char) seq[L - 1 - p]] : lut.clsF[(unsigned char) seq[p]];
}

__device__ __forceinline__ char orf_char(const char *__restrict__ seq, unsigned L, unsigned p, bool reverse, const OrfLut &lut) {
    if (p >= L) return ORF_PAD;
    char ch = seq[reverse ? (L - 1 - p) : p];
    if (ch == 'u') ch = 't';
    if (reverse) { ch = lut.rc[(unsigned char) ch]; if (ch == '.') ch = 'N'; }
    return ch;
}
__device__ __forceinline__ char orf_upper(char ch) { return ch == ORF_PAD ? ORF_PAD : (char) (ch & (char) ~0x20); }
__device__ __forceinline__ bool orf_is(char a, char b, char c, char x, char y, char z) { return a == x && b == y && c == z; }
__device__ __forceinline__ bool orf_gap(char ch, const OrfLut &lut) { return ch == 'N' || lut.rc[(unsigned char) ch] == '.'; }

struct OrfEmit {
    char *data;                        // output data file
    unsigned long long byteOff;        // running offset of this read's fragments
    unsigned long long *offsets;
    unsigned *lens, *keys, *info;      // info: 4 words per fragment
    unsigned long long index;          // running fragment index (= new key)
    unsigned readKey;
    const char *aminoAcid;             // m_AminoAcid[4097] in shared memory
};

template <bool EMIT>
__device__ void orf_scan_strand(const char *__restrict__ seq, unsigned L, bool reverse, unsigned frames, const OrfConst &c,
                                const char *__restrict__ aminoAcid, const OrfLut &lut, unsigned &nOrf, unsigned long long &nBytes, OrfEmit &e) {
    // Orf::findForward: per-frame state, initially inside an ORF that starts at the frame offset (:207-218)
    bool inside[3] = {true, true, true}, hasStart[3] = {false, false, false};
    unsigned gaps[3] = {0, 0, 0}, len[3] = {0, 0, 0}, from[3] = {0, 1, 2};
    const unsigned nPos = ((L - 2 + 2) / 3) * 3;                 // positions visited: i = 0, 3, ... < L - 2, position = i .. i + 2
    unsigned c0 = orf_class(seq, L, 0, reverse, lut), c1 = orf_class(seq, L, 1, reverse, lut), c2 = 7u;
    // nPos is a multiple of three: unrolling by the frame keeps the per-frame state in registers
    for (unsigned pos0 = 0; pos0 < nPos; pos0 += 3) {
#pragma unroll
      for (int frame = 0; frame < 3; frame++) {
        const unsigned position = pos0 + frame;
        if (position > 0) { c0 = c1; c1 = c2; }
        c2 = orf_class(seq, L, position + 2, reverse, lut);
        if (!(frames & (1u << frame))) continue;
        const unsigned codon = c0 * 64u + c1 * 8u + c2;
        const bool thisIncomplete = c0 == 7u || c1 == 7u || c2 == 7u;
        const bool isLast = !thisIncomplete && (position + 5 >= L);                       // the next codon of the frame runs into the padding
        const bool start = codon == ORF_ATG || (c.allStarts && (codon == ORF_TTG || codon == ORF_CTG));
        bool shouldStart;
        if (c.startMode == 0) shouldStart = !inside[frame] && start;
        else if (c.startMode == 1) shouldStart = !inside[frame];
        else shouldStart = start;
        if (shouldStart) { inside[frame] = true; hasStart[frame] = true; from[frame] = position; gaps[frame] = 0; len[frame] = 0; }
        const bool stop = codon == ORF_TAA || codon == ORF_TAG || codon == ORF_TGA;
        if (inside[frame]) {
            if (!stop) len[frame]++;
            // the padding counts as a gap as well: its complement is '.' (isGapOrN on CHAR_MAX)
            if (c0 == 5u || c1 == 5u || c2 == 5u || thisIncomplete) gaps[frame]++;
        }
        if (inside[frame] && (stop || isLast)) {
            inside[frame] = false;
            if (len[frame] == 0 && stop) continue;
            const unsigned to = (isLast && !stop) ? position + 2 : position - 1;
            if ((int) gaps[frame] > c.maxGaps || (int) len[frame] > c.maxLength || (int) len[frame] < c.minLength) continue;
            const bool incStart = !hasStart[frame], incEnd = !stop;
            if (c.contigStartMode < 2 && (int) incStart == c.contigStartMode) continue;   // extractorfs.cpp:85-90
            if (c.contigEndMode < 2 && (int) incEnd == c.contigEndMode) continue;
            const unsigned f = from[frame];
            const unsigned nt = to - f + 1;                                                // a multiple of three
            unsigned entryLen;
            bool addStart = false, addEnd = false;
            if (!c.translate) {
                entryLen = nt + 2;
            } else {
                // translatenucs.cpp:55-118: '*' in front of a complete start, '*' behind a complete end unless the last
                // residue already is one (ambiguity codes such as TAR translate to '*' without being a stop codon)
                addStart = !incStart; addEnd = !incEnd;
                if (addEnd) {
                    const unsigned i0 = lut.b2i[(unsigned char) orf_char(seq, L, to - 2, reverse, lut)], i1 = lut.b2i[(unsigned char) orf_char(seq, L, to - 1, reverse, lut)],
                                   i2 = lut.b2i[(unsigned char) orf_char(seq, L, to, reverse, lut)];
                    if (aminoAcid[256 * i0 + 16 * i1 + i2 + 1] == '*') addEnd = false;
                }
                entryLen = nt / 3 + 1 + (addStart ? 1u : 0u) + (addEnd ? 1u : 0u) + 1;
            }
            if (EMIT) {
                char *w = e.data + e.byteOff;
                if (!c.translate) {
                    for (unsigned k = 0; k < nt; k++) w[k] = orf_char(seq, L, f + k, reverse, lut);
                    w[nt] = '\n'; w[nt + 1] = '\0';
                } else {
                    unsigned o = 0;
                    if (addStart) w[o++] = '*';
                    for (unsigned k = 0; k < nt; k += 3) {
                        const char a = orf_char(seq, L, f + k, reverse, lut), b = orf_char(seq, L, f + k + 1, reverse, lut), d = orf_char(seq, L, f + k + 2, reverse, lut);
                        const bool lower = (a >= 'a' && a <= 'z') || (b >= 'a' && b <= 'z') || (d >= 'a' && d <= 'z');
                        char res = aminoAcid[256 * lut.b2i[(unsigned char) a] + 16 * lut.b2i[(unsigned char) b] + lut.b2i[(unsigned char) d] + 1];
                        if (lower && res >= 'A' && res <= 'Z') res = (char) (res + 32);
                        w[o++] = res;
                    }
                    if (addEnd) w[o++] = '*';
                    w[o++] = '\n'; w[o] = '\0';
                }
                e.offsets[e.index] = e.byteOff;
                e.lens[e.index] = entryLen;
                e.keys[e.index] = (unsigned) e.index;
                unsigned fromPos = f, toPos = to;
                if (reverse) { fromPos = (L - 1) - f; toPos = (L - 1) - to; }                // extractorfs.cpp:95-100
                e.info[4 * e.index + 0] = e.readKey; e.info[4 * e.index + 1] = fromPos; e.info[4 * e.index + 2] = toPos;
                e.info[4 * e.index + 3] = (incStart ? 1u : 0u) | (incEnd ? 2u : 0u);
                e.byteOff += entryLen;
                e.index++;
            }
            nOrf++;
            nBytes += entryLen;
        }
      }
    }
}

// (Copying every read into a per-thread slot of shared memory before walking it was measured: 41 ms instead of 35 ms for
// 5 M reads -- the 46 KB of slots cut the occupancy to 16 warps / SM, which costs more than the byte-wise global reads.)

template <bool EMIT>
__global__ void __launch_bounds__(128) orf_kernel(const pg_seqdb db, const OrfConst c, const char *__restrict__ aminoAcidG,
                                                  unsigned *__restrict__ cnt, unsigned *__restrict__ bytes,
                                                  const unsigned long long *__restrict__ cntOff, const unsigned long long *__restrict__ byteOff,
                                                  char *__restrict__ oData, unsigned long long *__restrict__ oOffsets, unsigned *__restrict__ oLens,
                                                  unsigned *__restrict__ oKeys, unsigned *__restrict__ oInfo) {
    extern __shared__ __align__(16) unsigned char orf_smem[];
    char *sAmino = reinterpret_cast<char *>(orf_smem);                    // 4112 bytes
    char *sRc = sAmino + 4112;                                            // 256
    unsigned char *sB2i = reinterpret_cast<unsigned char *>(sRc + 256);   // 256
    unsigned char *sClsF = sB2i + 256, *sClsR = sB2i + 512;               // 2 x 256
    for (int i = threadIdx.x; i < 4097; i += blockDim.x) sAmino[i] = aminoAcidG[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { sRc[i] = c_orf_rc[i]; sB2i[i] = c_orf_b2i[i]; sClsF[i] = c_orf_clsF[i]; sClsR[i] = c_orf_clsR[i]; }
    __syncthreads();
    OrfLut lut; lut.rc = sRc; lut.b2i = sB2i; lut.clsF = sClsF; lut.clsR = sClsR;
    const unsigned n = (unsigned) db.n;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned L = db.lens[i] - 2;
        unsigned nOrf = 0; unsigned long long nBytes = 0;
        OrfEmit e;
        if (EMIT) {
            if (cnt[i] == 0) continue;
            e.data = oData; e.byteOff = byteOff[i]; e.offsets = oOffsets; e.lens = oLens; e.keys = oKeys; e.info = oInfo;
            e.index = cntOff[i]; e.readKey = db.keys[i]; e.aminoAcid = sAmino;
        }
        if (L >= 3) {                                                                      // Orf::setSequence (Orf.cpp:127-131)
            const char *seq = db.data + db.offsets[i];
            if (c.forwardFrames) orf_scan_strand<EMIT>(seq, L, false, c.forwardFrames, c, sAmino, lut, nOrf, nBytes, e);
            if (c.reverseFrames) orf_scan_strand<EMIT>(seq, L, true, c.reverseFrames, c, sAmino, lut, nOrf, nBytes, e);
        }
        if (!EMIT) { cnt[i] = nOrf; bytes[i] = (unsigned) nBytes; }
    }
}

// ---- translatenucs (lib/mmseqs/src/util/translatenucs.cpp:14-128) on an arbitrary nucleotide DB -----------------------------
// flags[i]: bit 0 = put '*' in front (the ORF header says "complete start"), bit 1 = put '*' behind unless the last residue
// is one already.  The length arithmetic follows the reference literally: it works on the entry INCLUDING its '\n'
// (length = entryLen - 1), trims only when residues % 3 == 1, and lets the residue at index length / 3 be overwritten by
// the terminator, so residues % 3 == 2 yields one extra 'X' (the codon that contains the '\n').
__device__ __forceinline__ bool tn_lengths(unsigned entryLen, unsigned &length, unsigned &nAa) {
    length = entryLen - 1;                                   // residues + '\n'
    const unsigned residues = length - 1;
    if (length % 3 != 0 && residues % 3 != 0) length -= length % 3;      // :70-73 (data[length] is the '\0', data[length-1] the '\n')
    if (length < 3) return false;                            // :75-78 entry skipped
    nAa = length / 3;
    return true;
}

template <bool EMIT>
__global__ void __launch_bounds__(128) tn_kernel(const pg_seqdb db, const unsigned char *__restrict__ flags, const char *__restrict__ aminoAcidG,
                                                 unsigned *__restrict__ keep, unsigned *__restrict__ outLen,
                                                 const unsigned long long *__restrict__ keepIdx, const unsigned long long *__restrict__ outOff,
                                                 char *__restrict__ oData, unsigned long long *__restrict__ oOffsets, unsigned *__restrict__ oLens, unsigned *__restrict__ oKeys) {
    __shared__ char sAmino[4112];
    __shared__ unsigned char sB2i[256];
    for (int i = threadIdx.x; i < 4097; i += blockDim.x) sAmino[i] = aminoAcidG[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sB2i[i] = c_orf_b2i[i];
    __syncthreads();
    const unsigned n = (unsigned) db.n;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned length, nAa;
        const bool ok = tn_lengths(db.lens[i], length, nAa);
        if (!ok) { if (!EMIT) { keep[i] = 0; outLen[i] = 0; } continue; }
        const char *seq = db.data + db.offsets[i];
        const unsigned fl = flags ? flags[i] : 0u;
        const bool addStart = fl & 1u;
        bool addEnd = (fl & 2u) != 0;
        auto residue = [&](unsigned k) {                     // codon k; bytes beyond the residues are the entry's own "\n\0"
            const char a = seq[3 * k], b = seq[3 * k + 1], d = seq[3 * k + 2];
            const bool lower = (a >= 'a' && a <= 'z') || (b >= 'a' && b <= 'z') || (d >= 'a' && d <= 'z');
            char res = sAmino[256 * sB2i[(unsigned char) a] + 16 * sB2i[(unsigned char) b] + sB2i[(unsigned char) d] + 1];
            if (lower && res >= 'A' && res <= 'Z') res = (char) (res + 32);
            return res;
        };
        if (addEnd && residue(nAa - 1) == '*') addEnd = false;
        const unsigned entryLen = nAa + 1 + (addStart ? 1u : 0u) + (addEnd ? 1u : 0u) + 1;
        if (!EMIT) { keep[i] = 1; outLen[i] = entryLen; continue; }
        char *w = oData + outOff[i];
        unsigned o = 0;
        if (addStart) w[o++] = '*';
        for (unsigned k = 0; k < nAa; k++) w[o++] = residue(k);
        if (addEnd) w[o++] = '*';
        w[o++] = '\n'; w[o] = '\0';
        const unsigned long long j = keepIdx[i];
        oOffsets[j] = outOff[i]; oLens[j] = entryLen; oKeys[j] = db.keys[i];
    }
}

// TranslateNucl(CANONICAL): sm_BaseToIdx and m_AminoAcid (TranslateNucl.h:333-470), built on the host once
static void orf_build_tables(unsigned char *baseToIdx, char *aminoAcid) {
    static const char charToBase[17] = "-ACMGRSVTWYHKDBN";
    memset(baseToIdx, 0, 256);
    for (int i = 0; i <= 15; i++) {
        baseToIdx[(unsigned char) charToBase[i]] = (unsigned char) i;
        const char lc = (charToBase[i] >= 'A' && charToBase[i] <= 'Z') ? (char) (charToBase[i] + 32) : charToBase[i];
        baseToIdx[(unsigned char) lc] = (unsigned char) i;
    }
    baseToIdx['U'] = 8; baseToIdx['u'] = 8; baseToIdx['X'] = 15; baseToIdx['x'] = 15;
    for (int i = 0; i <= 15; i++) baseToIdx[i] = (unsigned char) i;
    const char *ncbieaa = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    const int expansions[4] = {1, 2, 4, 8};                       // A C G T
    const int codonIdx[9] = {0, 2, 1, 0, 3, 0, 0, 0, 0};          // T = 0, C = 1, A = 2, G = 3
    for (int i = 0; i <= 4096; i++) aminoAcid[i] = 'X';
    int st = 1;
    for (int i = 0; i <= 15; i++) for (int j = 0; j <= 15; j++) for (int k = 0; k <= 15; k++, st++) {
        char aa = 0;
        bool go = true;
        for (int p = 0; p < 4 && go; p++) { if (!(expansions[p] & i)) continue;
            for (int q = 0; q < 4 && go; q++) { if (!(expansions[q] & j)) continue;
                for (int r = 0; r < 4 && go; r++) { if (!(expansions[r] & k)) continue;
                    const char ch = ncbieaa[16 * codonIdx[expansions[p]] + 4 * codonIdx[expansions[q]] + codonIdx[expansions[r]]];
                    if (aa == 0) aa = ch;
                    else if (aa != ch) {
                        if ((aa == 'B' || aa == 'D' || aa == 'N') && (ch == 'D' || ch == 'N')) aa = 'B';
                        else if ((aa == 'Z' || aa == 'E' || aa == 'Q') && (ch == 'E' || ch == 'Q')) aa = 'Z';
                        else if ((aa == 'J' || aa == 'I' || aa == 'L') && (ch == 'I' || ch == 'L')) aa = 'J';
                        else aa = 'X';
                    }
                    if (aa == 'X') go = false;
                } } }
        if (aa != 0) aminoAcid[st] = aa;
    }
}

static int orf_upload_tables(Context *ctx, char **dAmino, unsigned char *workspace) {
    static const char rcTable[257] =
        "................................................................"
        ".TVGH..CD..M.KN...YSAABW.R.......tvgh..cd..m.kn...ysaabw.r......"
        "................................................................"
        "................................................................";
    static unsigned char b2i[256];
    static char amino[4100];
    orf_build_tables(b2i, amino);
    static unsigned char clsF[256], clsR[256];
    for (int ch = 0; ch < 256; ch++) {
        auto classify = [&](unsigned char u) -> unsigned char {          // u: already upper-cased (& ~0x20)
            if (u == 'A') return 0; if (u == 'C') return 1; if (u == 'G') return 2; if (u == 'T') return 3;
            if (u == 'N' || rcTable[u] == '.') return 5;
            return 4;
        };
        const unsigned char f = (ch == 'u') ? (unsigned char) 't' : (unsigned char) ch;       // Orf::setSequence (Orf.cpp:144-147)
        clsF[ch] = (f == 127) ? 7 : classify((unsigned char) (f & 0xDF));
        char r = rcTable[f];
        if (r == '.') r = 'N';
        clsR[ch] = classify((unsigned char) ((unsigned char) r & 0xDF));
    }
    PG_CUDA(cudaMemcpyToSymbolAsync(c_orf_clsF, clsF, 256, 0, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(cudaMemcpyToSymbolAsync(c_orf_clsR, clsR, 256, 0, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(cudaMemcpyToSymbolAsync(c_orf_rc, rcTable, 256, 0, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(cudaMemcpyToSymbolAsync(c_orf_b2i, b2i, 256, 0, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(cudaMemcpyAsync(workspace, amino, 4097, cudaMemcpyHostToDevice, ctx->stream));
    *dAmino = (char *) workspace;
    return 0;
}

int tn_run(Context *ctx, const pg_seqdb *db, const unsigned char *d_flags, int translationTable, pg_seqdb **outDb) {
    cudaStream_t s = ctx->stream;
    PG_CHECK(db->dbtype == PG_DBTYPE_NUCLEOTIDES, "translatenucs: nucleotide sequence DB expected");
    PG_CHECK(translationTable == 1, "translatenucs: only --translation-table 1 (canonical) is built on the GPU path");
    const unsigned n = (unsigned) db->n;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t) 15; return r; };
    const size_t oAmino = take(4112), oKeep = take(sizeof(unsigned) * ((size_t) n + 1)), oLen = take(sizeof(unsigned) * ((size_t) n + 1));
    const size_t oKeepIdx = take(sizeof(unsigned long long) * ((size_t) n + 2)), oOff = take(sizeof(unsigned long long) * ((size_t) n + 2));
    const size_t oScan = take(scan_workspace_bytes(n));
    PG_TRY(ctx->nextWork.reserve(o));
    PG_TRY(ctx->small.reserve(4096));
    unsigned char *bb = ctx->nextWork.as<unsigned char>();
    char *dAmino = nullptr;
    PG_TRY(orf_upload_tables(ctx, &dAmino, bb + oAmino));
    unsigned *keep = (unsigned *) (bb + oKeep), *outLen = (unsigned *) (bb + oLen);
    unsigned long long *keepIdx = (unsigned long long *) (bb + oKeepIdx), *outOff = (unsigned long long *) (bb + oOff);
    unsigned long long totals[2] = {0, 0};
    const unsigned blocks = std::max(1u, std::min<unsigned>((n + 127) / 128, NUM_SMS * 16));
    if (n) {
        tn_kernel<false><<<blocks, 128, 0, s>>>(*db, d_flags, dAmino, keep, outLen, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        ctx->launches++;
        unsigned long long *d_tot = ctx->small.as<unsigned long long>() + 46;
        PG_TRY(exclusive_scan_u32(keep, keepIdx, n, d_tot, bb + oScan, scan_workspace_bytes(n), s, &ctx->launches));
        PG_TRY(exclusive_scan_u32(outLen, outOff, n, d_tot + 1, bb + oScan, scan_workspace_bytes(n), s, &ctx->launches));
        PG_TRY(read_back(ctx, totals, d_tot, sizeof(totals)));
    }
    pg_seqdb *out = new pg_seqdb();
    out->n = totals[0]; out->data_bytes = totals[1]; out->dbtype = PG_DBTYPE_AMINO_ACIDS;
    PG_CUDA(cudaMallocAsync(&out->data, totals[1] + 16, s));
    PG_CUDA(cudaMallocAsync(&out->offsets, sizeof(unsigned long long) * (totals[0] + 1), s));
    PG_CUDA(cudaMallocAsync(&out->lens, sizeof(unsigned) * (totals[0] + 1), s));
    PG_CUDA(cudaMallocAsync(&out->keys, sizeof(unsigned) * (totals[0] + 1), s));
    if (totals[0]) {
        tn_kernel<true><<<blocks, 128, 0, s>>>(*db, d_flags, dAmino, keep, outLen, keepIdx, outOff, out->data, out->offsets, out->lens, out->keys);
        ctx->launches++;
    }
    PG_CUDA(cudaGetLastError());
    PG_TRY(seqdb_finalize(ctx, out));
    *outDb = out;
    return 0;
}

int orf_run(Context *ctx, const pg_seqdb *db, const pg_orf_params *p, int translate, pg_seqdb **outDb, unsigned **d_info) {
    cudaStream_t s = ctx->stream;
    PG_CHECK(db->dbtype == PG_DBTYPE_NUCLEOTIDES, "extractorfs: nucleotide sequence DB expected");
    PG_CHECK(p->translation_table == 1, "extractorfs: only --translation-table 1 (canonical) is built on the GPU path");
    PG_CHECK(!(p->orf_start_mode == 1 && p->contig_start_mode < 2), "Parameter combination is illegal, orf-start-mode 1 can only go with contig-start-mode 2");
    PG_CHECK(p->orf_start_mode >= 0 && p->orf_start_mode <= 2, "extractorfs: --orf-start-mode must be 0, 1 or 2");
    const unsigned n = (unsigned) db->n;
    OrfConst c;
    c.minLength = p->min_length; c.maxLength = p->max_length; c.maxGaps = p->max_gaps;
    c.contigStartMode = p->contig_start_mode; c.contigEndMode = p->contig_end_mode; c.startMode = p->orf_start_mode;
    c.forwardFrames = p->forward_frames & 7u; c.reverseFrames = p->reverse_frames & 7u;
    c.allStarts = p->use_all_table_starts != 0; c.translate = translate != 0;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t) 15; return r; };
    const size_t oAmino = take(4112), oCnt = take(sizeof(unsigned) * ((size_t) n + 1)), oBytes = take(sizeof(unsigned) * ((size_t) n + 1));
    const size_t oCntOff = take(sizeof(unsigned long long) * ((size_t) n + 2)), oByteOff = take(sizeof(unsigned long long) * ((size_t) n + 2));
    const size_t oScan = take(scan_workspace_bytes(n));
    PG_TRY(ctx->nextWork.reserve(o));
    PG_TRY(ctx->small.reserve(4096));
    unsigned char *bb = ctx->nextWork.as<unsigned char>();
    char *dAmino = nullptr;
    PG_TRY(orf_upload_tables(ctx, &dAmino, bb + oAmino));
    unsigned *cnt = (unsigned *) (bb + oCnt), *bytes = (unsigned *) (bb + oBytes);
    unsigned long long *cntOff = (unsigned long long *) (bb + oCntOff), *byteOff = (unsigned long long *) (bb + oByteOff);
    unsigned long long totals[2] = {0, 0};
    const unsigned blocks = std::max(1u, std::min<unsigned>((n + 127) / 128, NUM_SMS * 16));
    constexpr size_t ORF_SMEM_BYTES = 4112 + 1024;
    {
        static std::atomic<unsigned long long> attrDev{0};
        if (first_use_on_device(attrDev)) {
            PG_CUDA(cudaFuncSetAttribute(orf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ORF_SMEM_BYTES));
            PG_CUDA(cudaFuncSetAttribute(orf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ORF_SMEM_BYTES));
        }
    }
    if (n) {
        orf_kernel<false><<<blocks, 128, ORF_SMEM_BYTES, s>>>(*db, c, dAmino, cnt, bytes, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        ctx->launches++;
        unsigned long long *d_tot = ctx->small.as<unsigned long long>() + 46;
        PG_TRY(exclusive_scan_u32(cnt, cntOff, n, d_tot, bb + oScan, scan_workspace_bytes(n), s, &ctx->launches));
        PG_TRY(exclusive_scan_u32(bytes, byteOff, n, d_tot + 1, bb + oScan, scan_workspace_bytes(n), s, &ctx->launches));
        PG_TRY(read_back(ctx, totals, d_tot, sizeof(totals)));
    }
    const unsigned long long nFrag = totals[0], nBytes = totals[1];
    PG_CHECK(nFrag < 0xFFFFFFF0ull, "extractorfs: more than 2^32 fragments");
    pg_seqdb *out = new pg_seqdb();
    out->n = nFrag; out->data_bytes = nBytes; out->dbtype = translate ? PG_DBTYPE_AMINO_ACIDS : PG_DBTYPE_NUCLEOTIDES;
    PG_CUDA(cudaMallocAsync(&out->data, nBytes + 16, s));
    PG_CUDA(cudaMallocAsync(&out->offsets, sizeof(unsigned long long) * (nFrag + 1), s));
    PG_CUDA(cudaMallocAsync(&out->lens, sizeof(unsigned) * (nFrag + 1), s));
    PG_CUDA(cudaMallocAsync(&out->keys, sizeof(unsigned) * (nFrag + 1), s));
    PG_TRY(ctx->orfInfo.reserve(sizeof(unsigned) * 4 * (nFrag + 1)));
    if (nFrag) {
        orf_kernel<true><<<blocks, 128, ORF_SMEM_BYTES, s>>>(*db, c, dAmino, cnt, bytes, cntOff, byteOff, out->data, out->offsets, out->lens, out->keys, ctx->orfInfo.as<unsigned>());
        ctx->launches++;
    }
    PG_CUDA(cudaGetLastError());
    PG_TRY(seqdb_finalize(ctx, out));
    *outDb = out;
    if (d_info) *d_info = ctx->orfInfo.as<unsigned>();
    return 0;
}

}  // namespace pg
