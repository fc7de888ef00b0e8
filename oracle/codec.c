/*
 * codec.c -- CPU ORACLE (test infrastructure, see vg_oracle.h) for the area save-file format.
 *
 * Follows service/persistence/world_persistence.rs:27-42,62-71 (store_blocking / load_blocking),
 * generic_persistence.rs:17-103 (read/write_binary_object with compression) and model/area.rs:204-227
 * (AreaDTO { voxels: Box<[Voxel]> }): a file is
 *     u32 LE uncompressed size  ||  LZ4 block of  bincode(AreaDTO)
 * with bincode 2.0.1 config::standard() (little endian, variable-length integers):
 *     slice length 32768 as varint = FB 00 80, then one byte per voxel (enum variant index < 251).
 *
 * The two crates are not under /root/reference (lz4_flex 0.11.5, bincode 2.0.1); what is restated
 * here are their published formats: the LZ4 block format and bincode's varint rule.  The
 * *compressor* below is an ordinary greedy hash-chain-free LZ4 compressor, not lz4_flex's: two
 * LZ4 compressors need not produce the same bytes, only streams that every LZ4 decoder expands
 * to the same data.  Parity for this format is therefore defined on the decoded bytes:
 * "compressed-bytes parity UNPINNED", container layout and decode results pinned by the format.
 */
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "vg_oracle.h"

#define PAYLOAD (3 + VGO_VOXELS_IN_AREA)

/* bincode: u64 varint of the slice length, then the variant index of every voxel */
void vgo_bincode_area(const uint8_t* voxels, uint8_t* out /* 32771 */) {
    out[0] = 251; /* U16_BYTE marker: 251 <= 32768 < 65536 */
    out[1] = (uint8_t)(VGO_VOXELS_IN_AREA & 0xFF);
    out[2] = (uint8_t)(VGO_VOXELS_IN_AREA >> 8);
    memcpy(out + 3, voxels, VGO_VOXELS_IN_AREA);
}

/* LZ4 block decoder.  Returns the number of bytes written or -1 on a malformed stream. */
long vgo_lz4_decompress_block(const uint8_t* src, long n, uint8_t* dst, long cap) {
    long ip = 0, op = 0;
    while (ip < n) {
        const unsigned token = src[ip++];
        long lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (ip >= n) return -1;
                b = src[ip++];
                lit += b;
            } while (b == 255);
        }
        if (ip + lit > n || op + lit > cap) return -1;
        memcpy(dst + op, src + ip, (size_t)lit);
        ip += lit;
        op += lit;
        if (ip == n) break; /* last sequence: literals only */
        if (ip + 2 > n) return -1;
        const long offset = src[ip] | (src[ip + 1] << 8);
        ip += 2;
        if (offset == 0 || offset > op) return -1;
        long len = (token & 15) + 4;
        if ((token & 15) == 15) {
            unsigned b;
            do {
                if (ip >= n) return -1;
                b = src[ip++];
                len += b;
            } while (b == 255);
        }
        if (op + len > cap) return -1;
        for (long k = 0; k < len; ++k) dst[op + k] = dst[op - offset + k]; /* overlap allowed */
        op += len;
    }
    return op;
}

static long put_length(uint8_t* dst, long op, long v) { /* the 255-run extension bytes of a length >= 15 */
    v -= 15;
    while (v >= 255) {
        dst[op++] = 255;
        v -= 255;
    }
    dst[op++] = (uint8_t)v;
    return op;
}

/* Greedy LZ4 block compressor with a 4-byte hash table (byte-granular matches, any offset up to
 * 65535, overlapping matches): stands in for "a file some LZ4 compressor wrote".  Returns the
 * compressed size (dst must hold n + n/255 + 16 bytes). */
long vgo_lz4_compress_block(const uint8_t* src, long n, uint8_t* dst) {
    enum { HASH_BITS = 13 };
    static _Thread_local int32_t table[1 << HASH_BITS];
    for (int i = 0; i < (1 << HASH_BITS); ++i) table[i] = -1;
    long ip = 0, anchor = 0, op = 0;
    const long match_limit = n - 12; /* the last match starts at least 12 bytes before the end */
    while (ip < match_limit) {
        uint32_t seq;
        memcpy(&seq, src + ip, 4);
        const uint32_t h = (seq * 2654435761u) >> (32 - HASH_BITS);
        const long cand = table[h];
        table[h] = (int32_t)ip;
        if (cand >= 0 && ip - cand <= 65535 && memcmp(src + cand, src + ip, 4) == 0) {
            long len = 4;
            while (ip + len < n - 5 && src[cand + len] == src[ip + len]) ++len; /* last 5 bytes are literals */
            const long lit = ip - anchor;
            uint8_t* token = &dst[op++];
            *token = (uint8_t)((lit >= 15 ? 15 : lit) << 4);
            if (lit >= 15) op = put_length(dst, op, lit);
            memcpy(dst + op, src + anchor, (size_t)lit);
            op += lit;
            dst[op++] = (uint8_t)((ip - cand) & 0xFF);
            dst[op++] = (uint8_t)((ip - cand) >> 8);
            const long ml = len - 4;
            *token |= (uint8_t)(ml >= 15 ? 15 : ml);
            if (ml >= 15) op = put_length(dst, op, ml);
            ip += len;
            anchor = ip;
        } else {
            ++ip;
        }
    }
    const long lit = n - anchor;
    dst[op++] = (uint8_t)((lit >= 15 ? 15 : lit) << 4);
    if (lit >= 15) op = put_length(dst, op, lit);
    memcpy(dst + op, src + anchor, (size_t)lit);
    return op + lit;
}

/* write_binary_object(&AreaDTO, compressed): returns the file size (out must hold 33 300 bytes) */
long vgo_area_file_encode(const uint8_t* voxels, uint8_t* out) {
    uint8_t payload[PAYLOAD];
    vgo_bincode_area(voxels, payload);
    const uint32_t size = PAYLOAD; /* compress_prepend_size: u32 LE */
    out[0] = (uint8_t)size;
    out[1] = (uint8_t)(size >> 8);
    out[2] = (uint8_t)(size >> 16);
    out[3] = (uint8_t)(size >> 24);
    return 4 + vgo_lz4_compress_block(payload, PAYLOAD, out + 4);
}

/* store_all_blocking (world_persistence.rs:54-58): every area through write_binary_object on the
 * rayon pool; out has n slots of VGO_AREA_FILE_MAX_BYTES, sizes[i] = bytes of file i. Returns the sum. */
long vgo_area_files_encode(const uint8_t* voxels, int n, int threads, uint8_t* out, long* sizes) {
    long total = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads) reduction(+ : total)
#endif
    for (int i = 0; i < n; ++i) {
        sizes[i] = vgo_area_file_encode(voxels + (size_t)i * VGO_VOXELS_IN_AREA, out + (size_t)i * VGO_AREA_FILE_MAX_BYTES);
        total += sizes[i];
    }
    (void)threads;
    return total;
}

/* read_binary_object::<AreaDTO>(compressed): 0 and the 32768 voxels, or -1 where the reference
 * would log an error and fall back to generation (world_persistence.rs:66-70): bad size prefix,
 * malformed LZ4 stream, wrong slice length, variant index >= 27. */
int vgo_area_file_decode(const uint8_t* file, long n, uint8_t* voxels) {
    if (n < 4) return -1;
    const uint32_t size = file[0] | (file[1] << 8) | (file[2] << 16) | ((uint32_t)file[3] << 24);
    if (size != PAYLOAD) return -1;
    uint8_t payload[PAYLOAD];
    if (vgo_lz4_decompress_block(file + 4, n - 4, payload, PAYLOAD) != PAYLOAD) return -1;
    if (payload[0] != 251 || payload[1] != (VGO_VOXELS_IN_AREA & 0xFF) || payload[2] != (VGO_VOXELS_IN_AREA >> 8)) return -1;
    for (int i = 0; i < VGO_VOXELS_IN_AREA; ++i)
        if (payload[3 + i] >= VGO_VOXEL_VARIANTS) return -1;
    memcpy(voxels, payload + 3, VGO_VOXELS_IN_AREA);
    return 0;
}
