// kq_csv.cu — CsvDataSource.scan (Main.kt:276-357) + ReaderIterator.createBatch (Main.kt:251-273) on the GPU:
// the text of a CSV file -> one RecordBatch of Utf8 columns resident in HBM, ready for the operators.
//
// The reference delegates tokenising to univocity-parsers (absent from /root/reference, version unpinned) with
// delimiter and line-separator detection, skipEmptyLines and header extraction (Main.kt:289-296, 322); every value
// then goes through getValue(name, "").trim() into a VarCharVector (Main.kt:262-264). Rules restated here (C1-C9,
// oracle: ko_csv_scan in oracle/kq_oracle.cpp; parity UNPINNED — see DESIGN.md):
//   C1 line separator: "\n" (a preceding "\r" is whitespace and trimmed) or, when the text holds no "\n", "\r";
//      delimiter: the most frequent of , ; TAB | outside quotes in the first record (ties: that order; the first record
//      is found with all four acting as delimiters).
//   C2 quote '"': a field whose first non-blank byte is a quote is quoted — delimiters and line separators inside are
//      data, "" is a literal quote, a single '"' closes, bytes between the closing quote and the next delimiter are
//      dropped. A quote anywhere else (inside an unquoted value, behind a closed section) is data, as CSV parsers treat it.
//   C3 records without any byte are skipped (skipEmptyLines).   C4 has_headers: the first record names the columns.
//   C5 each value is trimmed of leading/trailing bytes <= 0x20 (String.trim(), Main.kt:263), quoted or not.
//   C6 every column is Utf8 and never null: a missing or empty field reads "" (getValue's default, Main.kt:263).
//   C7 projection selects/reorders file columns (Main.kt:313-318).   C8 surplus fields of a record are ignored.
//   C9 one output batch per call (the reference cuts 1000-row batches, Main.kt:396; results do not depend on it).
//
// Device pipeline (all HBM-bound byte work; no tensor cores). Passes 1-3 see the text as one bit per byte (SWAR masks
// over 64-byte blocks), so no pass runs a per-byte state machine:
//   1. quote states    : every block's transition over the three quote states (a walk over the block's quotes; most blocks
//                        have none), composed per chunk of blocks, chained, expanded -> the quote state in front of every block
//   2. record scan,    : device-wide exclusive sums of (a) terminators outside quotes that end a non-empty record and
//      separator scan    (b) field separators = those terminators + delimiters outside quotes
//   3. separators      : the position of every separator in text order + each record's last separator index; field c of
//                        a record is then the text between two consecutive separators — no re-parsing
//   4. field lengths   : one thread per record, trimmed/unescaped length of every projected field
//   5. offsets         : one device-wide exclusive sum per projected column -> Arrow int32 offsets
//   6. copy            : one thread per record writes the field bytes
// kq_csv_reader_* runs the same passes piece by piece over texts of any size (Sequence<RecordBatch>, Main.kt:239-249).
// The file is also compiled for the HOST by the CPU test-suite (tests/test_csv_host.py, tests/host_shim/): keep the
// kernels free of warp-level intrinsics so that they stay checkable against the oracle without a GPU.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kq_internal.h"
#include "kq_scan.cuh"

using namespace kq;

namespace {

constexpr int CSV_BLOCK = 64;            // bytes per scan item
constexpr int CSV_MAX_COLS = 256;        // file columns

struct CsvFormat { uint8_t delim, term; };

__device__ __forceinline__ bool csv_blank(uint8_t c) { return c <= 0x20; }

// ---- rule C2 on the device ---------------------------------------------------------------------------------------
// Between two bytes the scan is outside quotes (Q_OUT), inside a quoted section (Q_IN), or outside right behind the quote
// that closed one (Q_OUTE: a '"' now is the second half of a doubled quote and re-opens the section). Only quotes change
// the state: IN -"-> OUTE, OUTE -"-> IN, OUT -"-> IN if the quote is the first non-blank byte of its field, else OUT (the
// quote is data); any other byte takes OUTE to OUT. Whether a quote stands at the start of a field is a LOCAL property
// once we know we are outside: its nearest non-blank byte in front is a delimiter or terminator (which, being outside as
// well, is a real one) or the text starts there. So a 64-byte block is one transition over three states, computed by
// walking the block's quotes (most blocks have none), and the state in front of every block is a prefix composition of
// those transitions (k_csv_quote_maps -> k_csv_compose_chunks -> k_csv_chunk_states -> k_csv_block_states).
enum : uint32_t { Q_OUT = 0, Q_IN = 1, Q_OUTE = 2 };
#ifndef KQ_CSV_CHUNK
#define KQ_CSV_CHUNK 1024                // the host build of the test-suite also runs with 16, so that small texts span many chunks
#endif
constexpr int CSV_CHUNK = KQ_CSV_CHUNK;  // blocks whose transitions one thread composes
static_assert(CSV_CHUNK % 16 == 0, "chunks start 16-byte aligned in the per-block state array");

__device__ __forceinline__ bool csv_pad(uint8_t c, CsvFormat f) { return c <= 0x20 && c != f.delim && c != f.term; }   // blank between a separator and a value

// One bit per byte of block i (bytes past the end of the text: zero)
struct RawMasks { uint64_t q, t, c, d; int nbytes; };
__device__ __forceinline__ uint64_t swar_zero4(uint32_t x) {          // bit k = byte k of x is zero
    x = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;        // 0x80 in every zero byte (exact)
    return (uint64_t)((((x >> 7) * 0x00204081u) >> 21) & 0xFu);
}
__device__ __forceinline__ RawMasks csv_raw_masks(const uint8_t* __restrict__ text, long long n, long long i, CsvFormat f) {
    const long long b = i * CSV_BLOCK;
    RawMasks m{0, 0, 0, 0, CSV_BLOCK};
    if (b + CSV_BLOCK > n) {
        m.nbytes = (int)(n - b);
        for (long long p = b; p < n; p++) {
            const uint8_t c = text[p];
            const uint64_t bit = 1ULL << (p - b);
            if (c == '"') m.q |= bit;
            if (c == f.term) m.t |= bit;
            if (c == '\r') m.c |= bit;
            if (c == f.delim) m.d |= bit;
        }
        return m;
    }
    const uint4* p4 = reinterpret_cast<const uint4*>(text + b);
    const uint32_t tpat = f.term * 0x01010101u, dpat = f.delim * 0x01010101u;
#pragma unroll
    for (int j = 0; j < CSV_BLOCK / 16; j++) {
        const uint4 v = __ldg(p4 + j);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int sh = 16 * j + 4 * t;
            m.q |= swar_zero4(w[t] ^ 0x22222222u) << sh;
            m.t |= swar_zero4(w[t] ^ tpat) << sh;
            m.c |= swar_zero4(w[t] ^ 0x0D0D0D0Du) << sh;
            m.d |= swar_zero4(w[t] ^ dpat) << sh;
        }
    }
    return m;
}
__device__ __forceinline__ bool csv_has_quote(const uint8_t* __restrict__ text, long long n, long long i) {
    const long long b = i * CSV_BLOCK;
    if (b + CSV_BLOCK > n) {
        for (long long p = b; p < n; p++) if (text[p] == '"') return true;
        return false;
    }
    const uint4* p4 = reinterpret_cast<const uint4*>(text + b);
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < CSV_BLOCK / 16; j++) {
        const uint4 v = __ldg(p4 + j);
        const uint32_t w[4] = {v.x ^ 0x22222222u, v.y ^ 0x22222222u, v.z ^ 0x22222222u, v.w ^ 0x22222222u};
#pragma unroll
        for (int t = 0; t < 4; t++) any |= ~(((w[t] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w[t]) & 0x80808080u;
    }
    return any != 0;
}
// Quotes of block i that can open a quoted section (only evaluated for blocks that hold a quote): field starts — the byte
// behind every delimiter/terminator, and the block's first byte if the text in front of it ends in [separator][blanks] —
// are carried through runs of blanks by one addition (the carry ripples through the run and stops behind it).
__device__ __noinline__ uint64_t csv_opener_candidates(const uint8_t* __restrict__ text, long long n, long long i, const RawMasks m, CsvFormat f) {
    const long long b = i * CSV_BLOCK;
    uint64_t blank = 0;
    if (m.nbytes == CSV_BLOCK) {
        const uint4* p4 = reinterpret_cast<const uint4*>(text + b);
#pragma unroll
        for (int j = 0; j < CSV_BLOCK / 16; j++) {
            const uint4 v = __ldg(p4 + j);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                // 0x80 in every byte <= 0x20: high bit clear and the low seven bits do not carry when 0x5F is added
                const uint32_t le = ~(((w[t] & 0x7F7F7F7Fu) + 0x5F5F5F5Fu) | w[t]) & 0x80808080u;
                blank |= (uint64_t)((((le >> 7) * 0x00204081u) >> 21) & 0xFu) << (16 * j + 4 * t);
            }
        }
        blank &= ~(m.d | m.t);
    } else {
        for (int j = 0; j < m.nbytes; j++) blank |= (uint64_t)csv_pad(text[b + j], f) << j;
    }
    long long p = b - 1;
    while (p >= 0 && csv_pad(text[p], f)) p--;
    const uint64_t carry = p < 0 || text[p] == f.delim || text[p] == f.term;
    const uint64_t start = ((m.d | m.t) << 1) | carry;
    const uint64_t reach = ((blank + (start & blank)) ^ blank) | start;
    return m.q & reach;
}
// The state behind a block whose quotes are `q` (of which `cand` may open a section), entered in state s; *sq receives the
// quotes that switch between inside and outside.
__device__ __forceinline__ uint32_t csv_quote_walk(int nbytes, uint64_t q, uint64_t cand, uint32_t s, uint64_t* sq) {
    uint64_t toggles = 0;
    int prev = -1;
    while (q) {
        const int j = __ffsll((long long)q) - 1;
        q &= q - 1;
        if (s == Q_OUTE && j != prev + 1) s = Q_OUT;
        if (s == Q_IN) { s = Q_OUTE; toggles |= 1ULL << j; }
        else if (s == Q_OUTE || ((cand >> j) & 1ULL)) { s = Q_IN; toggles |= 1ULL << j; }
        prev = j;
    }
    if (s == Q_OUTE && prev != nbytes - 1) s = Q_OUT;
    if (sq) *sq = toggles;
    return s;
}
// The block's transition: csv_quote_walk for the three states in front of it at once (one pass over the quotes)
__device__ __forceinline__ uint32_t csv_quote_map(int nbytes, uint64_t q, uint64_t cand) {
    uint32_t s[3] = {Q_OUT, Q_IN, Q_OUTE};
    int prev = -1;
    while (q) {
        const int j = __ffsll((long long)q) - 1;
        q &= q - 1;
        const bool adjacent = j == prev + 1, opener = (cand >> j) & 1ULL;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint32_t x = s[k];
            if (x == Q_OUTE && !adjacent) x = Q_OUT;
            s[k] = x == Q_IN ? Q_OUTE : (x == Q_OUTE || opener) ? Q_IN : Q_OUT;
        }
        prev = j;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) if (s[k] == Q_OUTE && prev != nbytes - 1) s[k] = Q_OUT;
    return s[0] | (s[1] << 2) | (s[2] << 4);
}
// transitions are three 2-bit states packed into a byte: bits [2s, 2s+1] = the state behind the block when s is in front
__device__ __forceinline__ uint32_t qmap_apply(uint32_t map, uint32_t s) { return (map >> (2 * s)) & 3u; }
__device__ __forceinline__ uint32_t qmap_then(uint32_t first, uint32_t second) {
    return qmap_apply(second, qmap_apply(first, Q_OUT)) | (qmap_apply(second, qmap_apply(first, Q_IN)) << 2) | (qmap_apply(second, qmap_apply(first, Q_OUTE)) << 4);
}
constexpr uint32_t QMAP_ID = Q_OUT | (Q_IN << 2) | (Q_OUTE << 4);

// Passes over all blocks whose rare blocks (those that hold a quote) need much more work than the rest: a thread first does
// the cheap part of `batch` blocks, remembering which of them need the rest, then the expensive parts one after another —
// the lanes of a warp then work on their few expensive blocks together instead of each making the other 31 wait at every block.
template <class Light, class Heavy>
__device__ __forceinline__ void csv_for_blocks(long long nblocks, int batch, Light light, Heavy heavy) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < nblocks; i0 += stride * batch) {
        uint32_t todo = 0;
        for (int k = 0; k < batch; k++) {
            const long long i = i0 + k * stride;
            if (i < nblocks && light(i)) todo |= 1u << k;
        }
        while (todo) {
            const int k = __ffs((int)todo) - 1;
            todo &= todo - 1;
            heavy(i0 + k * stride);
        }
    }
}
// pass 1a: every block's transition
__global__ void k_csv_quote_maps(const uint8_t* __restrict__ text, long long n, long long nblocks, CsvFormat f, int batch, uint8_t* __restrict__ qstate) {
    csv_for_blocks(nblocks, batch,
        [&](long long i) {
            if (csv_has_quote(text, n, i)) return true;
            qstate[i] = (uint8_t)(Q_OUT | (Q_IN << 2) | (Q_OUT << 4));            // no quote in the block
            return false;
        },
        [&](long long i) {
            const RawMasks m = csv_raw_masks(text, n, i, f);
            qstate[i] = (uint8_t)csv_quote_map(m.nbytes, m.q, csv_opener_candidates(text, n, i, m, f));
        });
}
// pass 1b: one transition per chunk of CSV_CHUNK blocks (16 transitions per load; chunks start 16-byte aligned)
__global__ void k_csv_compose_chunks(const uint8_t* __restrict__ qstate, long long nblocks, long long nchunks, uint8_t* __restrict__ chunk_map) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nchunks; k += (long long)gridDim.x * blockDim.x) {
        const long long i1 = (k + 1) * CSV_CHUNK < nblocks ? (k + 1) * CSV_CHUNK : nblocks;
        long long i = k * CSV_CHUNK;
        uint32_t map = QMAP_ID;
        for (; i + 16 <= i1; i += 16) {
            const uint4 v = *reinterpret_cast<const uint4*>(qstate + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; t++)
#pragma unroll
                for (int j = 0; j < 4; j++) map = qmap_then(map, (w[t] >> (8 * j)) & 0xFFu);
        }
        for (; i < i1; i++) map = qmap_then(map, qstate[i]);
        chunk_map[k] = (uint8_t)map;
    }
}
// pass 1c (one thread): the state in front of every chunk, and behind the text. Separate input and output arrays: stores
// into the array being read made every load of this dependent chain miss (110 us per 6000 chunks when done in place).
__global__ void k_csv_chunk_states(const uint8_t* __restrict__ chunk_map, long long nchunks, uint8_t* __restrict__ chunk_state, unsigned long long* final_state) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t s = Q_OUT;
    long long k = 0;
#pragma unroll 4
    for (; k + 16 <= nchunks; k += 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(chunk_map + k));
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; t++) {
            uint32_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                o |= s << (8 * j);
                s = qmap_apply((w[t] >> (8 * j)) & 0xFFu, s);
            }
            w[t] = o;
        }
        uint4 r; r.x = w[0]; r.y = w[1]; r.z = w[2]; r.w = w[3];
        *reinterpret_cast<uint4*>(chunk_state + k) = r;
    }
    for (; k < nchunks; k++) {
        chunk_state[k] = (uint8_t)s;
        s = qmap_apply(chunk_map[k], s);
    }
    *final_state = s;
}
// pass 1d: the state in front of every block (replaces the block's transition in place)
__global__ void k_csv_block_states(uint8_t* __restrict__ qstate, long long nblocks, long long nchunks, const uint8_t* __restrict__ chunk_state) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nchunks; k += (long long)gridDim.x * blockDim.x) {
        const long long i1 = (k + 1) * CSV_CHUNK < nblocks ? (k + 1) * CSV_CHUNK : nblocks;
        long long i = k * CSV_CHUNK;
        uint32_t s = chunk_state[k];
        for (; i + 16 <= i1; i += 16) {
            const uint4 v = *reinterpret_cast<const uint4*>(qstate + i);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                uint32_t o = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    o |= s << (8 * j);
                    s = qmap_apply((w[t] >> (8 * j)) & 0xFFu, s);
                }
                w[t] = o;
            }
            uint4 r; r.x = w[0]; r.y = w[1]; r.z = w[2]; r.w = w[3];
            *reinterpret_cast<uint4*>(qstate + i) = r;
        }
        for (; i < i1; i++) {
            const uint32_t map = qstate[i];
            qstate[i] = (uint8_t)s;
            s = qmap_apply(map, s);
        }
    }
}

// One bit per byte of block i: `rec` = terminators outside quotes that end a NON-EMPTY record (rule C3), `delim` =
// delimiters outside quotes. Quote state by prefix XOR over the quotes that switch it; a terminator ends an empty record
// when the byte in front of it is a terminator too (or a CR behind a terminator), tested with shifted masks (carries: the
// two bytes in front of the block; the start of the text counts as a terminator at position -1).
struct alignas(16) BlockMasks { uint64_t rec, delim; };
__device__ __forceinline__ BlockMasks csv_block_masks(const uint8_t* __restrict__ text, long long b, const RawMasks m, uint64_t inside, CsvFormat f) {
    const bool t1 = b == 0 || text[b - 1] == f.term, t2 = b <= 1 || text[b - 2] == f.term;
    uint64_t empty = (m.t << 1) | (uint64_t)t1;
    if (f.term == '\n') {
        const uint64_t prev_c = (m.c << 1) | (uint64_t)(b > 0 && text[b - 1] == '\r');
        const uint64_t prev2_t = (m.t << 2) | ((uint64_t)t1 << 1) | (uint64_t)t2;
        empty |= prev_c & prev2_t;
    }
    BlockMasks r;
    r.rec = m.t & ~inside & ~empty;
    r.delim = m.d & ~inside;
    return r;
}
// pass 2a: the masks of every block, computed once and kept in HBM (16 bytes per 64-byte block) for the two scans and the
// separator pass (rebuilding them from the text in each of the three was 13 % slower)
__global__ void k_csv_store_masks(const uint8_t* __restrict__ text, long long n, long long nblocks, const uint8_t* __restrict__ qstate, CsvFormat f, int batch,
                                  BlockMasks* __restrict__ masks) {
    csv_for_blocks(nblocks, batch,
        [&](long long i) {
            const RawMasks m = csv_raw_masks(text, n, i, f);
            if (m.q) return true;
            masks[i] = csv_block_masks(text, i * CSV_BLOCK, m, qstate[i] == Q_IN ? ~0ULL : 0ULL, f);      // no quote: the whole block is inside or outside
            return false;
        },
        [&](long long i) {
            const RawMasks m = csv_raw_masks(text, n, i, f);
            const uint32_t s = qstate[i];
            uint64_t x = 0;
            csv_quote_walk(m.nbytes, m.q, csv_opener_candidates(text, n, i, m, f), s, &x);
            x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16; x ^= x << 32;       // prefix XOR over the quotes that switch the state
            masks[i] = csv_block_masks(text, i * CSV_BLOCK, m, s == Q_IN ? ~x : x, f);             // in-quote state at every non-quote byte
        });
}
struct RecordCount {          // records ending in block i
    const BlockMasks* masks;
    __device__ __forceinline__ int operator()(long long i) const { return __popcll(masks[i].rec); }
};
struct SeparatorCount {       // field separators in block i: delimiters + record ends
    const BlockMasks* masks;
    __device__ __forceinline__ int operator()(long long i) const { const BlockMasks m = masks[i]; return __popcll(m.rec | m.delim); }
};
// pass 3: the position of every separator, in text order, and for every record the index of its LAST separator (its end):
// field c of record r is the text between separators rec_last[r-1] + c and rec_last[r-1] + c + 1.
__global__ void k_csv_separators(const BlockMasks* __restrict__ masks, long long nblocks, const int32_t* __restrict__ recs_before, const int32_t* __restrict__ seps_before,
                                        int32_t* sep, int32_t* rec_last) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nblocks; i += (long long)gridDim.x * blockDim.x) {
        const BlockMasks m = masks[i];
        uint64_t all = m.rec | m.delim;
        int32_t k = seps_before[i], r = recs_before[i];
        while (all) {
            const int j = __ffsll((long long)all) - 1;
            sep[k] = (int32_t)(i * CSV_BLOCK + j);
            if ((m.rec >> j) & 1ULL) rec_last[r++] = k;
            k++;
            all &= all - 1;
        }
    }
}

// Trimmed, unquoted value of the raw field [first, last): returns its length; with `out`, also writes it.
__device__ __forceinline__ int csv_value(const uint8_t* text, long long first, long long last, uint8_t* out) {
    while (first < last && csv_blank(text[first])) first++;
    while (last > first && csv_blank(text[last - 1])) last--;
    if (first < last && text[first] == '"') {                  // quoted (rule C2): content up to the closing quote
        long long p = first + 1, q = p;
        // find the closing quote: a quote not followed by a quote
        while (q < last && !(text[q] == '"' && !(q + 1 < last && text[q + 1] == '"'))) q += (text[q] == '"') ? 2 : 1;
        if (q > last) q = last;
        // trim the content (String.trim() runs on the parsed value, rule C5)
        while (p < q && csv_blank(text[p])) p++;
        while (q > p && csv_blank(text[q - 1])) q--;
        int len = 0;
        for (long long j = p; j < q; j++) {
            if (out) out[len] = text[j];
            len++;
            if (text[j] == '"' && j + 1 < q && text[j + 1] == '"') j++;       // "" -> "
        }
        return len;
    }
    const int len = (int)(last - first);
    if (out) for (int j = 0; j < len; j++) out[j] = text[first + j];
    return len;
}

struct CsvCols {
    int16_t file_col[CSV_MAX_COLS];     // output column -> file column (materialised columns only; -1: shares another column's buffers)
    int32_t* lens[CSV_MAX_COLS];        // per output column: lengths, later Arrow offsets (device)
    uint8_t* data[CSV_MAX_COLS];
    int nout;
};

// Raw bytes [a, b) of file column c of record `rec` (a missing field: a == b, rule C6)
__device__ __forceinline__ void csv_field(const int32_t* __restrict__ sep, const int32_t* __restrict__ rec_last, long long rec, int c, long long& a, long long& b) {
    const long long k0 = rec ? (long long)rec_last[rec - 1] + 1 : 0, k1 = rec_last[rec];      // the record's separators: k0 .. k1
    a = b = 0;
    if (k0 + c > k1) return;
    a = k0 + c ? (long long)sep[k0 + c - 1] + 1 : 0;      // bytes of skipped empty lines in front of a record are blanks: trimmed with the value
    b = sep[k0 + c];
}
// pass 4: lengths of the projected fields of every data record; pass 6 (COPY): the bytes, at the offsets pass 5 produced.
// The column table travels as a kernel parameter (4.6 KB, read through the constant bank), not as a host-to-device copy.
template <bool COPY>
__global__ void k_csv_fields(const uint8_t* __restrict__ text, const int32_t* __restrict__ sep, const int32_t* __restrict__ rec_last, long long nrec, int skip,
                             const __grid_constant__ CsvCols cols) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r + skip < nrec; r += (long long)gridDim.x * blockDim.x) {
        for (int oc = 0; oc < cols.nout; oc++) {
            const int c = cols.file_col[oc];
            if (c < 0) continue;
            long long a, b;
            csv_field(sep, rec_last, r + skip, c, a, b);
            if (COPY) csv_value(text, a, b, cols.data[oc] + cols.lens[oc][r]);
            else cols.lens[oc][r] = csv_value(text, a, b, nullptr);
        }
    }
}
// Start of a device-wide scan: ticket, total and tile descriptors cleared, the item count in word 2 — for `nscans` scans
// whose scratch areas of `words` 64-bit words lie back to back. A kernel, not a memset plus a small host-to-device copy:
// such copies queue on the copy engine behind the reader's upload of the next piece and would serialise scan and upload.
__global__ void k_csv_scan_begin(unsigned long long* scratch, long long words, int nscans, unsigned long long count) {
    const long long total = words * nscans;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) scratch[i] = (i % words == 2) ? count : 0ULL;
}
// a reader's piece: the byte behind its last complete record (0: indices out of range)
__global__ void k_csv_last_end(const int32_t* __restrict__ sep, const int32_t* __restrict__ rec_last, long long nrec, long long nsep, unsigned long long* out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const long long k = rec_last[nrec - 1];
    *out = k >= 0 && k < nsep && sep[k] >= 0 ? (unsigned long long)sep[k] + 1ULL : 0ULL;
}
struct LenAt {
    const int32_t* lens;
    __device__ __forceinline__ int operator()(long long i) const { return lens[i]; }
};

// ---- host side: format detection and the header record (rules C1, C4) ---------------------------------------------
struct HostRecord { std::vector<std::string> fields; int64_t end = 0; };    // end: index just past the record's terminator

static void host_trim(std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= 0x20) a++;
    while (b > a && (unsigned char)s[b - 1] <= 0x20) b--;
    s = s.substr(a, b - a);
}
// The first NON-EMPTY record under rules C2 and C3, with the bytes of `delims` separating fields: its bytes [b, e) without
// the terminator and the positions of its delimiters outside quotes; false if the text holds no record. The same state
// machine as the device's (csv_quote_walk), one byte at a time; `fresh` = only blanks since the last separator.
static bool host_first_span(const uint8_t* t, int64_t n, const char* delims, uint8_t term, int64_t* b, int64_t* e, std::vector<int64_t>* seps) {
    int64_t start = 0;
    uint32_t state = Q_OUT;
    bool fresh = true;
    seps->clear();
    for (int64_t p = 0; p <= n; p++) {
        const bool at_end = p == n;
        const uint8_t c = at_end ? term : t[p];
        if (c == '"' && !at_end) {
            if (state == Q_IN) state = Q_OUTE;
            else if (state == Q_OUTE) state = Q_IN;
            else { if (fresh) state = Q_IN; fresh = false; }
            continue;
        }
        if (state == Q_OUTE) state = Q_OUT;
        if (state == Q_IN && !at_end) continue;
        if (c == term) {
            const bool empty = p == start || (term == '\n' && p == start + 1 && t[start] == '\r');
            if (!empty) { *b = start; *e = p; return true; }
            start = p + 1; fresh = true; seps->clear();
        } else if (c && strchr(delims, c)) { seps->push_back(p); fresh = true; }
        else if (c > 0x20) fresh = false;
    }
    return false;
}
static CsvFormat host_detect(const uint8_t* t, int64_t n) {
    CsvFormat f{',', '\n'};
    if (!memchr(t, '\n', (size_t)n) && memchr(t, '\r', (size_t)n)) f.term = '\r';
    // the first record is found with all four candidates acting as delimiters; the most frequent outside quotes wins
    int64_t b, e;
    std::vector<int64_t> seps;
    if (!host_first_span(t, n, ",;\t|", f.term, &b, &e, &seps)) return f;
    int64_t cnt[4] = {0, 0, 0, 0};
    const uint8_t cand[4] = {',', ';', '\t', '|'};
    for (int64_t p : seps) for (int k = 0; k < 4; k++) cnt[k] += t[p] == cand[k];
    int best = 0;
    for (int k = 1; k < 4; k++) if (cnt[k] > cnt[best]) best = k;
    f.delim = cand[best];
    return f;
}
static HostRecord host_first_record(const uint8_t* t, int64_t n, CsvFormat f) {
    HostRecord r;
    int64_t b, e;
    std::vector<int64_t> seps;
    const char delims[2] = {(char)f.delim, 0};
    if (!host_first_span(t, n, delims, f.term, &b, &e, &seps)) { r.end = n; return r; }
    r.end = e;
    seps.push_back(e);
    int64_t first = b;
    for (int64_t p : seps) {
        std::string raw((const char*)t + first, (size_t)(p - first));
        host_trim(raw);
        if (!raw.empty() && raw[0] == '"') {            // csv_value on the host
            std::string v;
            size_t q = 1;
            while (q < raw.size() && !(raw[q] == '"' && !(q + 1 < raw.size() && raw[q + 1] == '"'))) { v += raw[q]; q += raw[q] == '"' ? 2 : 1; }
            host_trim(v);
            raw = v;
        }
        r.fields.push_back(raw);
        first = p + 1;
    }
    return r;
}

// ---- what both entry points (kq_csv_scan, kq_csv_reader_*) do first: where the text lives, format, header, projection ----
struct CsvSource {
    bool on_device = false;
    uint8_t last_byte = 0;
    CsvFormat f{',', '\n'};
    int file_cols = 0;
    std::vector<int> proj;          // file column of every output column
};
static int csv_prologue(kq_ctx* ctx, const uint8_t* text, int64_t nbytes, const int* projection, int nproj, CsvSource* src) {
    // `text` may also be a DEVICE pointer (a file already resident in HBM: bench.py's device-resident timing). Format
    // detection and the header record then work on a copy of its first MiB (the first record must end inside it).
    src->on_device = false;
    {
        cudaPointerAttributes at;
        if (nbytes && cudaPointerGetAttributes(&at, text) == cudaSuccess) src->on_device = at.type == cudaMemoryTypeDevice;
        else cudaGetLastError();
    }
    std::vector<uint8_t> prefix;
    src->last_byte = nbytes && !src->on_device ? text[nbytes - 1] : 0;
    const uint8_t* htext = text;
    int64_t hbytes = nbytes;
    if (src->on_device) {
        hbytes = std::min<int64_t>(nbytes, 1 << 20);
        prefix.resize((size_t)hbytes);
        KQ_CUDA(ctx, cudaMemcpyAsync(prefix.data(), text, (size_t)hbytes, cudaMemcpyDeviceToHost, ctx->stream));
        KQ_CUDA(ctx, cudaMemcpyAsync(&src->last_byte, text + nbytes - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
        KQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        htext = prefix.data();
    }
    src->f = nbytes ? host_detect(htext, hbytes) : CsvFormat{',', '\n'};
    const HostRecord head = nbytes ? host_first_record(htext, hbytes, src->f) : HostRecord();
    if (src->on_device && hbytes < nbytes && head.end >= hbytes) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "first CSV record longer than 1 MiB");
    src->file_cols = (int)head.fields.size();
    if (src->file_cols > CSV_MAX_COLS) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d CSV columns", CSV_MAX_COLS);
    src->proj.clear();
    if (nproj) src->proj.assign(projection, projection + nproj);
    else for (int i = 0; i < src->file_cols; i++) src->proj.push_back(i);
    for (int c : src->proj)
        if (c < 0 || c >= src->file_cols) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "projected CSV column %d out of range (file has %d)", c, src->file_cols);   // Schema.select, Main.kt:47-52
    return KQ_OK;
}

// Passes 1-6 over `n` bytes of text resident in HBM (16-byte aligned): one batch of the projected columns.
// Whole text (partial = false): the text ends with a terminator; an unbalanced quote is an error.
// A reader's piece (partial = true): the text starts at a record start and may stop anywhere, also inside a quoted field.
// Every record end found is a true one (the quote state is exact from a record start on); *consumed = the byte after the
// last of them, the rest belongs to the next piece. A piece without any record end is an error.
static int csv_scan_resident(kq_ctx* ctx, const uint8_t* d_text, long long n, CsvFormat f, int file_cols, const std::vector<int>& proj, int has_headers,
                             bool partial, kq_batch** out, int64_t* consumed) {
    const int nout = (int)proj.size();
    const long long nblocks = (n + CSV_BLOCK - 1) / CSV_BLOCK;
    static const int batch = [] { const char* e = getenv("KQ_CSV_BATCH"); const int b = e ? atoi(e) : 16; return b < 1 ? 1 : b > 32 ? 32 : b; }();
    uint8_t *d_q = nullptr, *d_chunk = nullptr;     // per block: quote transition, then the quote state in front of it; per chunk likewise
    BlockMasks* d_masks = nullptr;                  // per block: record-end and delimiter bits
    int32_t* d_r = nullptr;
    unsigned long long* d_scratch = nullptr;        // two scan areas of [0] ticket, [1] unused, [2] item count, [4..] tile descriptors
    // what the host reads back, with as few synchronisations as there are decisions to take: [0] quote state behind the text,
    // [1] records, [2] separators, [3] end of the last complete record (a reader's piece), [8 + c] bytes of output column c
    unsigned long long* d_res = nullptr;
    int32_t *d_s = nullptr, *d_sep = nullptr, *d_last = nullptr;
    std::vector<kq_col*> cols;
    // tile descriptors of the device-wide scans: over 64-byte blocks (pass 2) and over records (pass 5; a record has at
    // least two bytes, so nblocks * 32 bounds the record count)
    const long long ntiles = (std::max<long long>(nblocks, 1) * (CSV_BLOCK / 2) + SCAN_TILE - 1) / SCAN_TILE + 2;
    auto cleanup = [&](int st) {
        kq_dev_free(ctx, d_q); kq_dev_free(ctx, d_chunk); kq_dev_free(ctx, d_masks); kq_dev_free(ctx, d_r); kq_dev_free(ctx, d_scratch); kq_dev_free(ctx, d_res); kq_dev_free(ctx, d_s); kq_dev_free(ctx, d_sep); kq_dev_free(ctx, d_last);
        if (st != KQ_OK) for (kq_col* c : cols) kq_column_free(c);
        return st;
    };
    auto make_batch = [&](int64_t rows) {
        kq_batch* b = new kq_batch();
        b->ctx = ctx; b->n = rows; b->cols = cols;
        *out = b;
        return KQ_OK;
    };
    int64_t nrec = 0, nsep = 0;
    int st = KQ_OK;
    if (consumed) *consumed = n;
    if (n > 0) {
        const long long nchunks = (nblocks + CSV_CHUNK - 1) / CSV_CHUNK;
        if ((st = kq_dev_alloc(ctx, (size_t)nblocks + 16, (void**)&d_q)) != KQ_OK) return cleanup(st);
        const size_t chunk_pitch = ((size_t)nchunks + 255) / 256 * 256;         // [transitions][states], both 16-byte aligned
        if ((st = kq_dev_alloc(ctx, 2 * chunk_pitch, (void**)&d_chunk)) != KQ_OK) return cleanup(st);
        if ((st = kq_dev_alloc(ctx, (size_t)(nblocks + 1) * 4, (void**)&d_r)) != KQ_OK) return cleanup(st);
        const long long words = ntiles + 4;
        if ((st = kq_dev_alloc(ctx, (size_t)words * 2 * 8, (void**)&d_scratch)) != KQ_OK) return cleanup(st);
        if ((st = kq_dev_alloc(ctx, (size_t)(8 + nout) * 8, (void**)&d_res)) != KQ_OK) return cleanup(st);
        if ((st = kq_dev_alloc(ctx, (size_t)(nblocks + 1) * 4, (void**)&d_s)) != KQ_OK) return cleanup(st);
        if ((st = kq_dev_alloc(ctx, (size_t)nblocks * sizeof(BlockMasks), (void**)&d_masks)) != KQ_OK) return cleanup(st);
        const int sg = (int)std::max<long long>(1, std::min<long long>((nblocks + SCAN_TILE - 1) / SCAN_TILE, (long long)ctx->sm_count * 4));
        // 1. the quote state in front of every block (rule C2): block transitions, composed per chunk, chained, expanded
        const int gq = (int)std::max<long long>(1, std::min<long long>((nblocks + 255) / 256, (long long)ctx->sm_count * 8));
        const int gc = (int)std::max<long long>(1, std::min<long long>((nchunks + 63) / 64, (long long)ctx->sm_count * 8));
        k_csv_quote_maps<<<gq, 256, 0, ctx->stream>>>(d_text, n, nblocks, f, batch, d_q);
        k_csv_compose_chunks<<<gc, 64, 0, ctx->stream>>>(d_q, nblocks, nchunks, d_chunk);
        k_csv_chunk_states<<<1, 32, 0, ctx->stream>>>(d_chunk, nchunks, d_chunk + chunk_pitch, d_res);
        k_csv_block_states<<<gc, 64, 0, ctx->stream>>>(d_q, nblocks, nchunks, d_chunk + chunk_pitch);
        // 2. the block masks, then records and separators before every block: two scans queued back to back, each with its own
        //    ticket and tile descriptors, their totals next to the quote state — one read-back for the three
        k_csv_store_masks<<<gq, 256, 0, ctx->stream>>>(d_text, n, nblocks, d_q, f, batch, d_masks);
        const int gb = (int)std::min<long long>((2 * words + 255) / 256, 1024);
        k_csv_scan_begin<<<gb, 256, 0, ctx->stream>>>(d_scratch, words, 2, (unsigned long long)nblocks);
        unsigned long long* sa = d_scratch;
        unsigned long long* sb = d_scratch + words;
        k_exclusive_offsets<RecordCount><<<sg, 256, 0, ctx->stream>>>(RecordCount{d_masks}, sa + 2, d_r, sa + 4, (unsigned int*)sa, d_res + 1);
        k_exclusive_offsets<SeparatorCount><<<sg, 256, 0, ctx->stream>>>(SeparatorCount{d_masks}, sb + 2, d_s, sb + 4, (unsigned int*)sb, d_res + 2);
        if (cudaGetLastError() != cudaSuccess) return cleanup(kq_cuda_fail(ctx, cudaGetLastError(), "CSV passes 1-2"));
        ctx->launches += 8;
        uint64_t res[3] = {0, 0, 0};
        if ((st = kq_read_u64(ctx, d_res, 3, res)) != KQ_OK) return cleanup(st);
        if (!partial && res[0] == Q_IN) return cleanup(kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "CSV text ends inside a quoted field"));
        nrec = (int64_t)res[1];
        nsep = (int64_t)res[2];
        if (partial && nrec == 0) return cleanup(kq_fail(ctx, KQ_ERR_UNSUPPORTED, "CSV record longer than the reader's piece (%lld bytes): open the reader with larger pieces", n));
    }
    const int skip = has_headers && nrec > 0 ? 1 : 0;
    const int64_t rows = nrec - skip;
    if (nrec > 0 && (partial || (rows > 0 && nout > 0))) {
        // 3. separator positions and each record's last separator
        if ((st = kq_dev_alloc(ctx, (size_t)nsep * 4 + 16, (void**)&d_sep)) != KQ_OK) return cleanup(st);
        if ((st = kq_dev_alloc(ctx, (size_t)nrec * 4 + 16, (void**)&d_last)) != KQ_OK) return cleanup(st);
        const int g = (int)std::max<long long>(1, std::min<long long>((nblocks + 255) / 256, (long long)ctx->sm_count * 8));
        k_csv_separators<<<g, 256, 0, ctx->stream>>>(d_masks, nblocks, d_r, d_s, d_sep, d_last);
        if (cudaGetLastError() != cudaSuccess) return cleanup(kq_cuda_fail(ctx, cudaGetLastError(), "k_csv_separators"));
        ctx->launches++;
        if (partial) {       // where the last complete record ends: separator rec_last[nrec - 1] (delimiters of the unfinished tail follow it)
            k_csv_last_end<<<1, 32, 0, ctx->stream>>>(d_sep, d_last, nrec, nsep, d_res + 3);
            ctx->launches++;
            uint64_t end = 0;
            if ((st = kq_read_u64(ctx, d_res + 3, 1, &end)) != KQ_OK) return cleanup(st);
            if (end == 0 || end > (uint64_t)n) return cleanup(kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "CSV separator index out of range"));
            *consumed = (int64_t)end;
        }
    }
    if (rows <= 0 || nout == 0) {            // no data records: zero-row columns (Main.kt:245-247 yields no batch; one empty batch here, rule R10's shape)
        for (int c = 0; c < nout; c++) {
            kq_col* col = nullptr;
            if ((st = kq_col_new(ctx, KQ_UTF8, 0, false, 0, &col)) != KQ_OK) return cleanup(st);
            cudaMemsetAsync(col->offsets, 0, 4, ctx->stream);
            cols.push_back(col);
        }
        make_batch(std::max<int64_t>(rows, 0));
        return cleanup(KQ_OK);
    }
    // 4. field lengths (into scratch arrays; the columns' offsets buffers are written by the scans below)
    CsvCols hc;
    memset(&hc, 0, sizeof hc);
    for (int i = 0; i < CSV_MAX_COLS; i++) hc.file_col[i] = -1;
    hc.nout = nout;
    std::vector<int32_t*> d_len((size_t)nout, nullptr);
    auto cleanup2 = [&](int s2) { for (int32_t* p : d_len) kq_dev_free(ctx, p); return cleanup(s2); };
    // a file column projected twice is materialised once and shared (ColumnExpression aliasing, rule R4)
    std::vector<int> first_out((size_t)file_cols, -1);
    for (int c = 0; c < nout; c++) {
        if (first_out[(size_t)proj[(size_t)c]] >= 0) continue;
        first_out[(size_t)proj[(size_t)c]] = c;
        hc.file_col[c] = (int16_t)proj[(size_t)c];
        if ((st = kq_dev_alloc(ctx, (size_t)(rows + 1) * 4, (void**)&d_len[(size_t)c])) != KQ_OK) return cleanup2(st);
        hc.lens[c] = d_len[(size_t)c];
    }
    const int gr = (int)std::max<long long>(1, std::min<long long>((rows + 127) / 128, (long long)ctx->sm_count * 16));
    k_csv_fields<false><<<gr, 128, 0, ctx->stream>>>(d_text, d_sep, d_last, nrec, skip, hc);
    if (cudaGetLastError() != cudaSuccess) return cleanup2(kq_cuda_fail(ctx, cudaGetLastError(), "k_csv_fields(lengths)"));
    ctx->launches++;
    // 5. offsets per materialised column, then the data buffers
    std::vector<uint64_t> bytes((size_t)nout, 0);
    std::vector<int32_t*> d_off((size_t)nout, nullptr);
    auto cleanup3 = [&](int s2) { for (int32_t* p : d_off) kq_dev_free(ctx, p); return cleanup2(s2); };
    const int sgr = (int)std::max<long long>(1, std::min<long long>((rows + SCAN_TILE - 1) / SCAN_TILE, (long long)ctx->sm_count * 4));
    // all column scans are queued back to back (each with its own ticket / total / tile descriptors), then the totals are read
    const size_t scratch_words = (size_t)(ntiles + 4);
    unsigned long long* d_scratch2 = nullptr;
    if ((st = kq_dev_alloc(ctx, scratch_words * 8 * (size_t)nout, (void**)&d_scratch2)) != KQ_OK) return cleanup3(st);
    auto cleanup4 = [&](int s2) { kq_dev_free(ctx, d_scratch2); return cleanup3(s2); };
    const unsigned long long count = (unsigned long long)rows;
    const int gb2 = (int)std::min<long long>(((long long)scratch_words * nout + 255) / 256, 1024);
    k_csv_scan_begin<<<gb2, 256, 0, ctx->stream>>>(d_scratch2, (long long)scratch_words, nout, count);
    ctx->launches++;
    for (int c = 0; c < nout; c++) {
        if (!d_len[(size_t)c]) continue;
        if ((st = kq_dev_alloc(ctx, (size_t)(rows + 1) * 4, (void**)&d_off[(size_t)c])) != KQ_OK) return cleanup4(st);
        unsigned long long* sc = d_scratch2 + scratch_words * (size_t)c;
        k_exclusive_offsets<LenAt><<<sgr, 256, 0, ctx->stream>>>(LenAt{d_len[(size_t)c]}, sc + 2, d_off[(size_t)c], sc + 4, (unsigned int*)sc, d_res + 8 + c);
        if (cudaGetLastError() != cudaSuccess) return cleanup4(kq_cuda_fail(ctx, cudaGetLastError(), "k_exclusive_offsets(column)"));
        ctx->launches++;
    }
    for (int c0 = 0; c0 < nout; c0 += 64)           // the totals of all column scans: one read-back per 64 columns
        if ((st = kq_read_u64(ctx, d_res + 8 + c0, std::min(64, nout - c0), bytes.data() + c0)) != KQ_OK) return cleanup4(st);
    for (int c = 0; c < nout; c++) {
        if (!d_len[(size_t)c]) { bytes[(size_t)c] = 0; continue; }       // shares another column's buffers: its slot was never written
        if (bytes[(size_t)c] >= (1ULL << 31)) return cleanup4(kq_fail(ctx, KQ_ERR_UNSUPPORTED, "CSV column of 2 GiB or more"));
    }
    kq_dev_free(ctx, d_scratch2);
    cols.assign((size_t)nout, nullptr);
    for (int c = 0; c < nout; c++) {
        if (!d_len[(size_t)c]) continue;
        kq_col* col = nullptr;
        if ((st = kq_col_new(ctx, KQ_UTF8, rows, false, (int64_t)bytes[(size_t)c], &col)) != KQ_OK) { cols.erase(std::remove(cols.begin(), cols.end(), nullptr), cols.end()); return cleanup3(st); }
        kq_dev_free(ctx, col->offsets);             // the scan's output IS the offsets buffer (same size and allocator): no copy
        col->offsets = d_off[(size_t)c];
        d_off[(size_t)c] = nullptr;
        cols[(size_t)c] = col;
        hc.lens[c] = col->offsets;
        hc.data[c] = (uint8_t*)col->data;
    }
    for (int c = 0; c < nout; c++)
        if (!cols[(size_t)c]) { kq_col* src = cols[(size_t)first_out[(size_t)proj[(size_t)c]]]; src->rc.fetch_add(1); cols[(size_t)c] = src; }
    // 6. the bytes (the table of pointers changed: lens now = offsets)
    k_csv_fields<true><<<gr, 128, 0, ctx->stream>>>(d_text, d_sep, d_last, nrec, skip, hc);
    st = cudaGetLastError() != cudaSuccess ? kq_cuda_fail(ctx, cudaGetLastError(), "k_csv_fields(copy)") : KQ_OK;
    ctx->launches++;
    if (st == KQ_OK) st = cudaStreamSynchronize(ctx->stream) == cudaSuccess ? KQ_OK : kq_cuda_fail(ctx, cudaGetLastError(), "cudaStreamSynchronize(csv)");
    if (st != KQ_OK) return cleanup3(st);
    make_batch(rows);
    return cleanup3(KQ_OK);
}

// ---- CsvDataSource.scan as a Sequence<RecordBatch> (Main.kt:239-249, 304-326): the text streams through two device
// buffers piece by piece; the H2D copy of piece k+1 (copy stream) runs under the scan of piece k (compute stream).
// A piece is cut where its last complete record ends; the unfinished tail is moved in front of the next piece.
// Buffer layout: [reserve R = piece][payload: piece][terminator + slack]; the text of a piece starts at `start` <= R
// (16-byte aligned: the bytes between `start` and the carried tail are terminators, i.e. empty lines, rule C3).
struct CsvPiece {
    uint8_t* buf = nullptr;
    int64_t start = 0;          // first byte of this piece's text in buf
    int64_t len = 0;            // payload bytes uploaded at buf + R
    bool last = false;          // the payload reaches the end of the text
    cudaEvent_t uploaded = nullptr;
};
}  // namespace

extern "C" {

int kq_csv_header(const uint8_t* text, int64_t nbytes, int has_headers, char* names, size_t names_cap, int* ncols, char* delimiter) {
    if (!text || nbytes < 0 || !ncols) return KQ_ERR_ILLEGAL_ARGUMENT;
    const CsvFormat f = host_detect(text, nbytes);
    const HostRecord h = host_first_record(text, nbytes, f);
    *ncols = (int)h.fields.size();
    if (delimiter) *delimiter = (char)f.delim;
    if (names && names_cap) {
        std::string all;
        for (size_t i = 0; i < h.fields.size(); i++) {
            all += has_headers ? h.fields[i] : "field_" + std::to_string(i + 1);      // Main.kt:345-349
            all += '\n';
        }
        if (all.size() + 1 > names_cap) return KQ_ERR_ILLEGAL_ARGUMENT;
        memcpy(names, all.c_str(), all.size() + 1);
    }
    return KQ_OK;
}

int kq_csv_scan(kq_ctx* ctx, const uint8_t* text, int64_t nbytes, int has_headers, const int* projection, int nproj, kq_batch** out) {
    if (!ctx || !out || nbytes < 0 || (nbytes && !text) || nproj < 0 || (nproj && !projection)) return KQ_ERR_ILLEGAL_ARGUMENT;
    if (nbytes >= (1LL << 31) - 2 * CSV_BLOCK) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "CSV text of 2 GiB or more: read it through kq_csv_reader_open (Arrow int32 offsets bound one batch)");
    cudaSetDevice(ctx->device);
    CsvSource src;
    KQ_RET(csv_prologue(ctx, text, nbytes, projection, nproj, &src));
    // the text on the device, terminated (a last record without a line separator still ends)
    const bool add_term = nbytes > 0 && src.last_byte != src.f.term;
    const long long n = nbytes + (add_term ? 1 : 0);
    // a text that is already resident, terminated and 16-byte aligned is scanned where it lies
    const bool in_place = src.on_device && !add_term && ((uintptr_t)text & 15u) == 0;
    uint8_t* d_text = nullptr;
    if (n > 0 && !in_place) {
        KQ_RET(kq_dev_alloc(ctx, (size_t)n + 16, (void**)&d_text));
        if (cudaMemcpyAsync(d_text, text, (size_t)nbytes, src.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            kq_dev_free(ctx, d_text);
            return kq_cuda_fail(ctx, cudaGetLastError(), "cudaMemcpyAsync(csv text)");
        }
        if (add_term) cudaMemsetAsync(d_text + nbytes, src.f.term, 1, ctx->stream);
    }
    const int st = csv_scan_resident(ctx, in_place ? text : d_text, n, src.f, src.file_cols, src.proj, has_headers, false, out, nullptr);
    kq_dev_free(ctx, d_text);
    return st;
}

struct kq_csv_reader {
    kq_ctx* ctx = nullptr;
    const uint8_t* text = nullptr;
    int64_t nbytes = 0;
    int has_headers = 0;
    CsvSource src;
    int64_t piece = 0;          // payload bytes per piece = reserve in front of it (multiple of 256)
    int64_t pos = 0;            // source bytes already handed to a copy
    CsvPiece p[2];
    int cur = 0;
    bool first = true, done = false;
};

static void csv_reader_upload(kq_csv_reader* r, int slot) {
    kq_ctx* ctx = r->ctx;
    CsvPiece& P = r->p[slot];
    P.len = std::min<int64_t>(r->piece, r->nbytes - r->pos);
    P.last = r->pos + P.len == r->nbytes;
    // the buffer may still be read by work queued on the compute stream (the previous scan of this slot, the copy of its tail)
    cudaEventRecord(ctx->copy_done, ctx->stream);
    cudaStreamWaitEvent(ctx->copy_stream[0], ctx->copy_done, 0);
    if (P.len) cudaMemcpyAsync(P.buf + r->piece, r->text + r->pos, (size_t)P.len, r->src.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->copy_stream[0]);
    cudaEventRecord(P.uploaded, ctx->copy_stream[0]);
    r->pos += P.len;
}

int kq_csv_reader_open(kq_ctx* ctx, const uint8_t* text, int64_t nbytes, int has_headers, const int* projection, int nproj,
                       int64_t piece_bytes, kq_csv_reader** out) {
    if (!ctx || !out || nbytes < 0 || (nbytes && !text) || nproj < 0 || (nproj && !projection) || piece_bytes < 0) return KQ_ERR_ILLEGAL_ARGUMENT;
    cudaSetDevice(ctx->device);
    kq_csv_reader* r = new kq_csv_reader();
    r->ctx = ctx; r->text = text; r->nbytes = nbytes; r->has_headers = has_headers;
    int st = csv_prologue(ctx, text, nbytes, projection, nproj, &r->src);
    if (st != KQ_OK) { delete r; return st; }
    // piece size: 64 MiB unless the caller says otherwise; never more than the text, at most 512 MiB (carry + piece stay below 2 GiB)
    int64_t piece = piece_bytes ? piece_bytes : (64LL << 20);
    piece = std::min<int64_t>(piece, 512LL << 20);
    piece = std::min<int64_t>(piece, std::max<int64_t>(nbytes, 1));
    r->piece = std::max<int64_t>(256, (piece + 255) / 256 * 256);
    for (int i = 0; i < 2 && st == KQ_OK; i++) {
        st = kq_dev_alloc(ctx, (size_t)(2 * r->piece + 64), (void**)&r->p[i].buf);
        if (st == KQ_OK && cudaEventCreateWithFlags(&r->p[i].uploaded, cudaEventDisableTiming) != cudaSuccess) st = kq_cuda_fail(ctx, cudaGetLastError(), "cudaEventCreate");
        r->p[i].start = r->piece;
    }
    if (st != KQ_OK) { kq_csv_reader_close(r); return st; }
    csv_reader_upload(r, 0);
    *out = r;
    return KQ_OK;
}

int kq_csv_reader_next(kq_csv_reader* r, kq_batch** out) {
    if (!r || !out) return KQ_ERR_ILLEGAL_ARGUMENT;
    kq_ctx* ctx = r->ctx;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    while (!r->done) {
        CsvPiece& P = r->p[r->cur];
        CsvPiece& N = r->p[r->cur ^ 1];
        const int64_t R = r->piece;
        cudaStreamWaitEvent(ctx->stream, P.uploaded, 0);
        if (!P.last) csv_reader_upload(r, r->cur ^ 1);          // runs under the scan below
        uint8_t* t = P.buf + P.start;
        long long n = (R - P.start) + P.len;
        if (P.last && r->nbytes > 0 && r->src.last_byte != r->src.f.term) {       // a last record without a line separator still ends
            cudaMemsetAsync(P.buf + R + P.len, r->src.f.term, 1, ctx->stream);
            n++;
        }
        kq_batch* b = nullptr;
        int64_t consumed = n;
        const int st = csv_scan_resident(ctx, t, n, r->src.f, r->src.file_cols, r->src.proj, r->first && r->has_headers, !P.last, &b, &consumed);
        if (st != KQ_OK) { r->done = true; return st; }
        r->first = false;
        if (P.last) r->done = true;
        else {
            // the unfinished tail goes in front of the next piece's payload; the gap down to a 16-byte boundary reads as empty lines
            const int64_t carry = n - consumed;
            if (carry > R - 16) { kq_batch_free(b); r->done = true; return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "CSV record longer than the reader's piece (%lld bytes): open the reader with larger pieces", (long long)R); }
            N.start = (R - carry) / 16 * 16;
            if (carry) cudaMemcpyAsync(N.buf + R - carry, t + consumed, (size_t)carry, cudaMemcpyDeviceToDevice, ctx->stream);
            if (R - carry > N.start) cudaMemsetAsync(N.buf + N.start, r->src.f.term, (size_t)(R - carry - N.start), ctx->stream);
        }
        r->cur ^= 1;
        if (b->n > 0) { *out = b; return KQ_OK; }      // Main.kt:245-247: a batch is yielded only when it has rows
        kq_batch_free(b);
    }
    return KQ_OK;
}

int kq_csv_reader_close(kq_csv_reader* r) {
    if (!r) return KQ_OK;
    kq_ctx* ctx = r->ctx;
    cudaSetDevice(ctx->device);
    // a copy may still be in flight from the caller's text (a reader closed early): the text must be reusable on return
    cudaStreamSynchronize(ctx->copy_stream[0]);
    cudaEventRecord(ctx->copy_done, ctx->copy_stream[0]);
    cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0);
    for (int i = 0; i < 2; i++) {
        kq_dev_free(ctx, r->p[i].buf);
        if (r->p[i].uploaded) cudaEventDestroy(r->p[i].uploaded);
    }
    delete r;
    return KQ_OK;
}

}  // extern "C"
