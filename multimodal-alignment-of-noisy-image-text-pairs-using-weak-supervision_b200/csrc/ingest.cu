// ingest.cu -- the lexical term sets of the chunk table, built on the GPU.
//
// The reference decides "term t occurs in chunk j" with a Python substring test per (image, chunk, term):
//     chunk_text_lower = text_chunk["text"].lower()
//     matching_terms = sum(1 for term in lexical_components if term in chunk_text_lower)
// (src/insert_clip_embeddings.py:149-150).  Here the whole table is matched in one pass: bit t of row j of the
// result = terms[t] is a substring of chunk j's (already lower-cased) text, which is the T-bit set the scoring
// kernels popcount (rescore.cu: term_hits).  Matching is on UTF-8 bytes; UTF-8 is self-synchronising, so a byte
// match of a valid pattern in a valid text is a code-point match, exactly Python's `in` on str.
//
// One warp per chunk.  Lane l looks at text positions l, l + 32, ...: the byte there selects a bucket of terms
// (terms grouped by first byte, index in shared memory), and each of those few terms is compared byte by byte.
// HBM-bound work in principle (each text byte is read from HBM once, L1 serves the re-reads of a comparison);
// in practice bounded by the byte comparisons.  Algorithmic bytes per chunk: its text length + 8 * term_words.
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace mma {

constexpr int kIngestThreads = 256;
constexpr int kIngestWarps = kIngestThreads / 32;
constexpr int kMaxTermWords = 64;  // T <= 4096 terms

struct TermTable {
    const uint8_t *bytes;        // concatenated terms
    const int32_t *off;          // [T + 1]
    const int32_t *bucket_start; // [257] terms grouped by first byte
    const int32_t *bucket_term;  // [T_nonempty] term ids, bucket by bucket, increasing id inside a bucket
};

__global__ void __launch_bounds__(kIngestThreads)
term_bitsets_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ text_off, int64_t m, TermTable tt,
                    int term_words, const uint32_t *__restrict__ always, uint64_t *__restrict__ bits)
{
    __shared__ int32_t s_bucket[257];
    __shared__ uint32_t s_bits[kIngestWarps][2 * kMaxTermWords];
    for (int b = threadIdx.x; b < 257; b += kIngestThreads) s_bucket[b] = tt.bucket_start[b];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *mine = s_bits[warp];
    const int words32 = 2 * term_words;
    for (int64_t j = blockIdx.x * (int64_t)kIngestWarps + warp; j < m; j += (int64_t)gridDim.x * kIngestWarps) {
        for (int w = lane; w < words32; w += 32) mine[w] = always[w];  // empty terms occur in every text
        __syncwarp();
        const int64_t t0 = text_off[j], len = text_off[j + 1] - t0;
        const uint8_t *s = text + t0;
        for (int64_t p = lane; p < len; p += 32) {
            const int c0 = s[p];
            const int b0 = s_bucket[c0], b1 = s_bucket[c0 + 1];
            for (int b = b0; b < b1; ++b) {
                const int t = tt.bucket_term[b];
                if (mine[t >> 5] & (1u << (t & 31))) continue;  // found earlier (a stale read only costs a re-match)
                const int o = tt.off[t], tl = tt.off[t + 1] - o;
                if (p + tl > len) continue;
                int q = 1;
                while (q < tl && s[p + q] == tt.bytes[o + q]) ++q;
                if (q == tl) atomicOr(&mine[t >> 5], 1u << (t & 31));
            }
        }
        __syncwarp();
        // little-endian: 32-bit word 2w is the low half of 64-bit word w
        uint32_t *dst = reinterpret_cast<uint32_t *>(bits + j * term_words);
        for (int w = lane; w < words32; w += 32) dst[w] = mine[w];
        __syncwarp();
    }
}

// Host side of the term table: offsets as int32, terms grouped by first byte (counting sort), the bit mask of
// empty terms.  The table is tiny (T terms); this is plumbing, the matching itself runs in the kernel above.
struct TermTableHost {  // (also declared in api.cu, its only user)
    std::vector<int32_t> off, bucket_start, bucket_term;
    std::vector<uint32_t> always;
};

int build_term_table(const uint8_t *terms, const int64_t *term_off, int n_terms, int term_words, TermTableHost *out)
{
    if (n_terms < 0 || term_words < 1 || term_words > kMaxTermWords || (int64_t)term_words * 64 < n_terms) return -1;
    if (n_terms > 0 && (term_off[0] != 0 || term_off[n_terms] > 0x7FFFFFFFll)) return -2;
    TermTableHost &h = *out;
    h.off.assign(n_terms + 1, 0);
    h.bucket_start.assign(257, 0);
    h.always.assign(2 * term_words, 0u);
    for (int t = 0; t < n_terms; ++t) {
        if (term_off[t + 1] < term_off[t]) return -2;
        h.off[t + 1] = (int32_t)term_off[t + 1];
        if (term_off[t + 1] == term_off[t]) h.always[t >> 5] |= 1u << (t & 31);
        else h.bucket_start[terms[term_off[t]] + 1]++;
    }
    for (int b = 0; b < 256; ++b) h.bucket_start[b + 1] += h.bucket_start[b];
    h.bucket_term.assign(std::max(1, h.bucket_start[256]), 0);
    std::vector<int32_t> fill(h.bucket_start.begin(), h.bucket_start.end() - 1);
    for (int t = 0; t < n_terms; ++t)
        if (term_off[t + 1] > term_off[t]) h.bucket_term[fill[terms[term_off[t]]]++] = t;
    return 0;
}

cudaError_t launch_term_bitsets(const uint8_t *text, const int64_t *text_off, int64_t m, const uint8_t *term_bytes,
                                const int32_t *term_off, const int32_t *bucket_start, const int32_t *bucket_term,
                                const uint32_t *always, int term_words, uint64_t *bits, cudaStream_t st)
{
    if (m == 0) return cudaSuccess;
    TermTable tt;
    tt.bytes = term_bytes; tt.off = term_off; tt.bucket_start = bucket_start; tt.bucket_term = bucket_term;
    int64_t grid = (m + kIngestWarps - 1) / kIngestWarps;
    if (grid > 148 * 8) grid = 148 * 8;
    term_bitsets_kernel<<<(unsigned)grid, kIngestThreads, 0, st>>>(text, text_off, m, tt, term_words, always, bits);
    return cudaGetLastError();
}

} // namespace mma
