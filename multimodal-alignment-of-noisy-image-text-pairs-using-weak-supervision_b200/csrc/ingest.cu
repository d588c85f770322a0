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
// One warp per chunk, lane l on text positions l, l + 32, ...  The three bytes at a position (two of them come
// from the neighbouring lanes by shuffle) are looked up in a hash table of the terms' first three bytes, held in
// shared memory: almost every position misses and costs one probe; a hit compares the rest of the few terms of
// that group byte by byte.  Two-byte terms sit in the same table under their two bytes (a second probe per
// position, no comparison needed on a hit); one-byte terms go through a first-byte bucket list.  Each text byte is
// read from HBM once.  Algorithmic bytes per chunk: its text length + 8 * term_words.
#include "common.cuh"
#include <cuda_fp16.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace mma {
MMA_CHECK_DECL

constexpr int kIngestThreads = 256;
constexpr int kIngestWarps = kIngestThreads / 32;
constexpr int kMaxTermWords = 64;  // T <= 4096 terms

struct TermTable {
    const uint8_t *bytes;        // concatenated terms
    const int32_t *off;          // [T + 1]
    const int32_t *bucket_start; // [257] one-byte terms, grouped by that byte
    const int32_t *bucket_term;  //       their ids, bucket by bucket
    const uint2 *hash;           // [hash_size] terms of >= 3 bytes by their first three bytes, x = 1 + (b0 | b1 << 8 |
                                 //   b2 << 16); two-byte terms by x = kTwoByteKey + (b0 | b1 << 8); 0 = empty slot;
                                 //   y = group start | count << 16
    const int32_t *group_term;   // term ids, group by group, increasing id inside a group
    int hash_bits;               // hash_size = 1 << hash_bits
};

constexpr uint32_t kTwoByteKey = 0x02000001u;
__host__ __device__ __forceinline__ uint32_t trigram_slot(uint32_t key, int bits)
{
    return (key * 2654435761u) >> (32 - bits);
}

__global__ void __launch_bounds__(kIngestThreads)
term_bitsets_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ text_off, int64_t m, TermTable tt,
                    int term_words, const uint32_t *__restrict__ always, uint64_t *__restrict__ bits)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2 *s_hash = reinterpret_cast<uint2 *>(smem_raw);
    __shared__ int32_t s_bucket[257];
    __shared__ uint32_t s_bits[kIngestWarps][2 * kMaxTermWords];
    const int hash_size = 1 << tt.hash_bits;
    for (int b = threadIdx.x; b < 257; b += kIngestThreads) s_bucket[b] = tt.bucket_start[b];
    for (int b = threadIdx.x; b < hash_size; b += kIngestThreads) s_hash[b] = tt.hash[b];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *mine = s_bits[warp];
    const int words32 = 2 * term_words;
    const uint32_t hmask = (uint32_t)hash_size - 1u;
    for (int64_t j = blockIdx.x * (int64_t)kIngestWarps + warp; j < m; j += (int64_t)gridDim.x * kIngestWarps) {
        for (int w = lane; w < words32; w += 32) mine[w] = always[w];  // empty terms occur in every text
        __syncwarp();
        const int64_t t0 = text_off[j], len = text_off[j + 1] - t0;
        const uint8_t *s = text + t0;
        for (int64_t base = 0; base < len; base += 32) {
            const int64_t p = base + lane;
            // this position's byte and the two after it: from the neighbouring lanes, and for the last two lanes
            // from the first two bytes of the next batch (fetched by lanes 0 and 1)
            const uint32_t c0 = p < len ? s[p] : 0u;
            const uint32_t nx = (lane < 2 && p + 32 < len) ? s[p + 32] : 0u;
            uint32_t c1 = __shfl_down_sync(0xFFFFFFFFu, c0, 1), c2 = __shfl_down_sync(0xFFFFFFFFu, c0, 2);
            const uint32_t n0 = __shfl_sync(0xFFFFFFFFu, nx, 0), n1 = __shfl_sync(0xFFFFFFFFu, nx, 1);
            if (lane == 31) { c1 = n0; c2 = n1; }
            if (lane == 30) c2 = n0;
            if (p >= len) continue;
            // one-byte terms
            for (int b = s_bucket[c0]; b < s_bucket[c0 + 1]; ++b) {
                const int t = tt.bucket_term[b];
                atomicOr(&mine[t >> 5], 1u << (t & 31));
            }
            // two-byte terms: the key is the whole term
            if (p + 2 > len) continue;
            {
                const uint32_t key2 = kTwoByteKey + (c0 | (c1 << 8));
                uint32_t h2 = trigram_slot(key2, tt.hash_bits);
                uint2 e2 = s_hash[h2];
                while (e2.x != 0u && e2.x != key2) { h2 = (h2 + 1u) & hmask; e2 = s_hash[h2]; }
                if (e2.x != 0u) {
                    const int a0 = (int)(e2.y & 0xFFFFu), a1 = a0 + (int)(e2.y >> 16);
                    for (int g = a0; g < a1; ++g) {
                        const int t = tt.group_term[g];
                        if (!(mine[t >> 5] & (1u << (t & 31)))) atomicOr(&mine[t >> 5], 1u << (t & 31));
                    }
                }
            }
            // terms of three bytes and more
            if (p + 3 > len) continue;
            const uint32_t key = 1u + (c0 | (c1 << 8) | (c2 << 16));
            uint32_t h = trigram_slot(key, tt.hash_bits);
            uint2 e = s_hash[h];
            while (e.x != 0u && e.x != key) { h = (h + 1u) & hmask; e = s_hash[h]; }
            if (e.x == 0u) continue;
            const int g0 = (int)(e.y & 0xFFFFu), g1 = g0 + (int)(e.y >> 16);
            for (int g = g0; g < g1; ++g) {
                const int t = tt.group_term[g];
                if (mine[t >> 5] & (1u << (t & 31))) continue;  // found earlier (a stale read only costs a re-match)
                const int o = tt.off[t], tl = tt.off[t + 1] - o;
                if (p + tl > len) continue;
                int q = 3;
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

// Host side of the term table: offsets as int32, one-byte terms grouped by their byte (counting sort), two-byte
// terms grouped by their two bytes and longer ones by their first three bytes behind one open-addressing hash
// table (load factor <= 1/2), the bit mask of empty terms.  The table is tiny (T terms); this is plumbing, the matching itself runs in the kernel above.
int build_term_table(const uint8_t *terms, const int64_t *term_off, int n_terms, int term_words, TermTableHost *out)
{
    if (n_terms < 0 || term_words < 1 || term_words > kMaxTermWords || (int64_t)term_words * 64 < n_terms) return -1;
    if (n_terms > 0 && (term_off[0] != 0 || term_off[n_terms] > 0x7FFFFFFFll)) return -2;
    TermTableHost &h = *out;
    h.off.assign(n_terms + 1, 0);
    h.bucket_start.assign(257, 0);
    h.always.assign(2 * term_words, 0u);
    std::vector<std::pair<uint32_t, int32_t>> longs;  // (hash key, term id)
    for (int t = 0; t < n_terms; ++t) {
        if (term_off[t + 1] < term_off[t]) return -2;
        h.off[t + 1] = (int32_t)term_off[t + 1];
        const int64_t tl = term_off[t + 1] - term_off[t];
        const uint8_t *b = terms + term_off[t];
        if (tl == 0) h.always[t >> 5] |= 1u << (t & 31);
        else if (tl == 1) h.bucket_start[b[0] + 1]++;
        else if (tl == 2) longs.push_back({kTwoByteKey + ((uint32_t)b[0] | ((uint32_t)b[1] << 8)), t});
        else longs.push_back({1u + ((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16)), t});
    }
    for (int b = 0; b < 256; ++b) h.bucket_start[b + 1] += h.bucket_start[b];
    h.bucket_term.assign(std::max(1, h.bucket_start[256]), 0);
    std::vector<int32_t> fill(h.bucket_start.begin(), h.bucket_start.end() - 1);
    for (int t = 0; t < n_terms; ++t) {
        const int64_t tl = term_off[t + 1] - term_off[t];
        if (tl == 1) h.bucket_term[fill[terms[term_off[t]]]++] = t;
    }
    std::sort(longs.begin(), longs.end());
    h.group_term.assign(std::max<size_t>(1, longs.size()), 0);
    size_t groups = 0;
    for (size_t i = 0; i < longs.size(); ++i) {
        h.group_term[i] = longs[i].second;
        if (i == 0 || longs[i].first != longs[i - 1].first) ++groups;
    }
    h.hash_bits = 6;
    while ((size_t)1 << h.hash_bits < 2 * groups) ++h.hash_bits;
    const uint32_t mask = (1u << h.hash_bits) - 1u;
    h.hash.assign((size_t)2 << h.hash_bits, 0u);
    for (size_t i = 0; i < longs.size();) {
        size_t e = i;
        while (e < longs.size() && longs[e].first == longs[i].first) ++e;
        const uint32_t key = longs[i].first;
        uint32_t slot = trigram_slot(key, h.hash_bits);
        while (h.hash[2 * slot] != 0u) slot = (slot + 1u) & mask;
        h.hash[2 * slot] = key;
        h.hash[2 * slot + 1] = (uint32_t)i | ((uint32_t)(e - i) << 16);
        i = e;
    }
    return 0;
}

cudaError_t launch_term_bitsets(const uint8_t *text, const int64_t *text_off, int64_t m, const uint8_t *term_bytes,
                                const int32_t *term_off, const int32_t *bucket_start, const int32_t *bucket_term,
                                const uint32_t *hash, int hash_bits, const int32_t *group_term,
                                const uint32_t *always, int term_words, uint64_t *bits, cudaStream_t st)
{
    if (m == 0) return cudaSuccess;
    TermTable tt;
    tt.bytes = term_bytes; tt.off = term_off; tt.bucket_start = bucket_start; tt.bucket_term = bucket_term;
    tt.hash = reinterpret_cast<const uint2 *>(hash); tt.group_term = group_term; tt.hash_bits = hash_bits;
    const size_t smem = (size_t)8 << hash_bits;  // <= 64 KiB (4096 groups)
    cudaError_t e = cudaFuncSetAttribute(term_bitsets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t grid = (m + kIngestWarps - 1) / kIngestWarps;
    if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
    term_bitsets_kernel<<<(unsigned)grid, kIngestThreads, smem, st>>>(text, text_off, m, tt, term_words, always, bits);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------
// pgvector interop: the `images` / `text_chunks` tables of
// src/setup_vector_db.py:102-131 arrive as a PostgreSQL binary COPY stream
//     COPY (SELECT image_id, manual_id, page, bbox, clip_embedding FROM <schema>.images ORDER BY id) TO STDOUT (FORMAT binary)
// Tuples are variable-length (VARCHAR ids, text), so their boundaries are found by one sequential walk on the host
// (copy_scan: a few loads per field); the bulk -- n x D big-endian float4 of the `vector` column (pgvector
// vector_send: int16 dim, int16 unused, float4[dim]), the REAL[] boxes and the INTEGER pages -- is decoded on the
// GPU straight into the layout mmalign_set_images / set_chunks take.
// ---------------------------------------------------------------------------------------------
static inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static inline uint16_t be16(const uint8_t *p) { return (uint16_t)(((uint16_t)p[0] << 8) | p[1]); }

// Walks the stream: field_off [rows][n_cols] = byte offset of each field's DATA, field_len = its length (-1 = NULL).
// Returns the number of tuples, or a negative code: -1 bad signature / flags, -2 truncated, -3 field count differs,
// -4 more tuples than `cap`.
int64_t copy_scan(const uint8_t *data, int64_t n_bytes, int n_cols, int64_t *field_off, int32_t *field_len, int64_t cap)
{
    static const uint8_t sig[11] = {'P', 'G', 'C', 'O', 'P', 'Y', '\n', 0xFF, '\r', '\n', 0};
    if (n_bytes < 19 || memcmp(data, sig, 11) != 0) return -1;
    if (be32(data + 11) & (1u << 16)) return -1;  // WITH OIDS
    int64_t o = 19 + (int64_t)be32(data + 15);
    int64_t rows = 0;
    for (;;) {
        if (o + 2 > n_bytes) return -2;
        const int16_t nf = (int16_t)be16(data + o);
        o += 2;
        if (nf == -1) return rows;
        if (nf != n_cols) return -3;
        if (rows >= cap && field_off) return -4;
        for (int f = 0; f < n_cols; ++f) {
            if (o + 4 > n_bytes) return -2;
            const int32_t ln = (int32_t)be32(data + o);
            o += 4;
            if (field_off) { field_off[rows * n_cols + f] = o; field_len[rows * n_cols + f] = ln; }
            if (ln > 0) o += ln;
            if (o > n_bytes) return -2;
        }
        ++rows;
    }
}

__device__ __forceinline__ uint32_t load_be32_unaligned(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8u;
    const uint32_t lo = w[0], hi = sh ? w[1] : 0u;          // little-endian words holding the four bytes
    return __byte_perm(__funnelshift_r(lo, hi, sh), 0u, 0x0123);  // bytes reversed: big-endian value
}

// One warp per tuple.  err: 1 = vector field malformed (NULL, wrong dim), set once; rows with a malformed box get zeros
// (corpus.bbox_array: missing / wrong-length boxes score 0.0, src/insert_clip_embeddings.py:161-169).
__global__ void __launch_bounds__(256)
copy_decode_kernel(const uint8_t *__restrict__ data, const int64_t *__restrict__ field_off, const int32_t *__restrict__ field_len,
                   int64_t n, int n_cols, int vec_col, int bbox_col, int page_col, int D, float *__restrict__ emb,
                   double *__restrict__ bbox, int32_t *__restrict__ page, uint8_t *__restrict__ page_null, int32_t *err)
{
    const int lane = threadIdx.x & 31;
    for (int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; r < n; r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t *off = field_off + r * n_cols;
        const int32_t *len = field_len + r * n_cols;
        if (emb) {
            const uint8_t *v = data + off[vec_col];
            bool ok = len[vec_col] == 4 + 4 * D;
            if (ok) ok = (load_be32_unaligned(v) == ((uint32_t)D << 16));  // int16 dim, int16 unused (0)
            if (!ok) { if (lane == 0) atomicExch(err, 1); }
            else
                for (int k = lane; k < D; k += 32) emb[r * D + k] = __uint_as_float(load_be32_unaligned(v + 4 + 4 * k));
        }
        if (bbox && lane < 4) {
            double x = 0.0;
            bool ok = false;
            if (bbox_col >= 0 && len[bbox_col] == 20 + 4 * 8) {  // ndim, has-null, oid, (size, lower bound), 4 x (len, float4)
                const uint8_t *b = data + off[bbox_col];
                ok = load_be32_unaligned(b) == 1u && load_be32_unaligned(b + 8) == 700u && load_be32_unaligned(b + 12) == 4u;
                if (ok) {
                    for (int q = 0; q < 4; ++q) ok = ok && load_be32_unaligned(b + 20 + 8 * q) == 4u;  // no NULL element
                    if (ok) x = (double)__uint_as_float(load_be32_unaligned(b + 24 + 8 * lane));
                }
            }
            bbox[r * 4 + lane] = ok ? x : 0.0;
        }
        if (page && lane == 0) {
            const bool null = page_col < 0 || len[page_col] != 4;
            page[r] = null ? 0 : (int32_t)load_be32_unaligned(data + off[page_col]);
            if (page_null) page_null[r] = null ? 1 : 0;
        }
    }
}

cudaError_t launch_copy_decode(const uint8_t *data, const int64_t *field_off, const int32_t *field_len, int64_t n, int n_cols,
                               int vec_col, int bbox_col, int page_col, int D, float *emb, double *bbox, int32_t *page,
                               uint8_t *page_null, int32_t *err, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    int64_t grid = (n * 32 + 255) / 256;
    if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
    copy_decode_kernel<<<(unsigned)grid, 256, 0, st>>>(data, field_off, field_len, n, n_cols, vec_col, bbox_col, page_col, D,
                                                       emb, bbox, page, page_null, err);
    return cudaGetLastError();
}

// Encoder output in half precision (fp16 / bf16 rows, e.g. an OpenCLIP batch on the device): widened to the fp32
// master rows the exact rescoring reads -- every 16-bit value is exactly representable, so nothing is lost.
template <typename T>
__global__ void widen_rows_kernel(const T *__restrict__ src, int64_t count, float *__restrict__ dst)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}

cudaError_t launch_widen_rows(const void *src, int dtype, int64_t count, float *dst, cudaStream_t st)
{
    if (count == 0) return cudaSuccess;
    int64_t grid = (count + 255) / 256;
    if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
    if (dtype == MMALIGN_F16) widen_rows_kernel<<<(unsigned)grid, 256, 0, st>>>(reinterpret_cast<const __half *>(src), count, dst);
    else widen_rows_kernel<<<(unsigned)grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16 *>(src), count, dst);
    return cudaGetLastError();
}

MMA_CHECK_READER(check_read_ingest)

} // namespace mma
