// vx_codec.cu -- area save files on the device (K8 encode, K9 decode).
//
// Replaces, for batches of resident areas, store_blocking / load_blocking
// (service/persistence/world_persistence.rs:27-42,62-71) = write/read_binary_object with
// compression (generic_persistence.rs:17-103) of AreaDTO { voxels: Box<[Voxel]> } (area.rs:204-227):
//     file = u32 LE 32771  ||  LZ4 block of  [FB 00 80 (bincode varint 32768)] [32768 voxel bytes]
// so that world pre-generation ships a few KB per area over PCIe instead of 32 KiB, and saved worlds
// are expanded where the voxels are needed.
//
// Encode is a *row-granular* LZ4 compressor made for this data (index = x + 16 y + 256 z): a
// 16-byte row either equals the row before it (match, offset 16: air, rock, flat ground), or the
// row one slab up (match, offset 256: vertical structures), or is sent as 16 literals.  Runs of
// rows of one kind are one LZ4 match / literal run, so all decisions are per row and independent:
// kinds, run boundaries and output positions come from block-wide scans, no serial parse.  The
// stream is a plain LZ4 block (last row always literal: the format wants the last 5 bytes literal
// and the last match to start 12 bytes before the end), readable by any LZ4 decoder; it is not the
// byte sequence lz4_flex would write, which no decoder requires.
// Decode accepts any valid LZ4 block (byte-granular, overlapping matches): one warp per area walks
// the sequences, the 32 lanes copy.
#include "vx_kernels.hpp"

namespace vx {
namespace {

constexpr int ROWS = VOXELS_IN_AREA / 16;
constexpr uint32_t PAYLOAD = 3 + VOXELS_IN_AREA;  // bincode(AreaDTO)
enum RowKind : uint8_t { LIT = 0, M16 = 1, M256 = 2 };

__device__ __forceinline__ uint32_t ext_bytes(uint32_t v) { return v >= 15u ? 1u + (v - 15u) / 255u : 0u; }

__device__ __forceinline__ uint8_t* put_ext(uint8_t* p, uint32_t v) {  // v >= 15
    v -= 15u;
    while (v >= 255u) {
        *p++ = 255;
        v -= 255u;
    }
    *p++ = (uint8_t)v;
    return p;
}

__device__ __forceinline__ bool same_row(const uint4 a, const uint4 b) {
    return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}

// The size pass stages the area's rows and classifies them; the kinds (one byte per row) go to a
// scratch buffer, so the emit pass starts from 2 KiB per area instead of 32 and reads only the
// literal rows again.  (Classifying straight from registers -- eight rows per thread, the row before
// and the row one slab up by warp shuffles, nothing staged -- was measured and lost: 0.388 -> 0.41 ms
// at 5, 6 or 8 CTAs per SM, profiles/r02_ab_codec.log.)
struct SizeSmem {
    uint4 rows[ROWS];
    uint8_t kind[ROWS];
    uint32_t warp_tmp[8];
};
struct EmitSmem {
    uint32_t seg_pos[ROWS];   // by segment END row: where the segment's bytes start in the block
    uint32_t lit_base[ROWS];  // by segment START row: where a literal segment's literals start
    uint16_t seg_start[ROWS];
    uint8_t kind[ROWS];
    uint32_t warp_tmp[8];
};

// inclusive max-scan / exclusive sum-scan over 2048 values, 8 consecutive per thread
__device__ __forceinline__ uint32_t block_carry_max(uint32_t mine, uint32_t* warp_tmp, int tid) {
    // returns the max over all threads before `tid` (0 if none)
    const int lane = tid & 31, wid = tid >> 5;
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc = max(inc, o);
    }
    if (lane == 31) warp_tmp[wid] = inc;
    __syncthreads();
    uint32_t before = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) before = 0;
    for (int w = 0; w < wid; ++w) before = max(before, warp_tmp[w]);
    __syncthreads();
    return before;
}

__device__ __forceinline__ uint32_t block_carry_sum(uint32_t mine, uint32_t* warp_tmp, int tid, uint32_t* total) {
    const int lane = tid & 31, wid = tid >> 5;
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_tmp[wid] = inc;
    __syncthreads();
    uint32_t before = inc - mine;
    uint32_t all = 0;
    for (int w = 0; w < 8; ++w) {
        if (w < wid) before += warp_tmp[w];
        all += warp_tmp[w];
    }
    *total = all;
    __syncthreads();
    return before;
}

// EMIT = false: sizes[i] = file size of area i, kinds[i] = its row kinds.
// EMIT = true: write the file at out + offsets[i].
template <bool EMIT>
__global__ void __launch_bounds__(256) codec_encode_kernel(const uint4* __restrict__ jobs,
                                                           const uint8_t* __restrict__ pool_voxels,
                                                           uint8_t* __restrict__ kinds,
                                                           uint32_t* __restrict__ sizes,
                                                           const unsigned long long* __restrict__ offsets,
                                                           uint8_t* __restrict__ out, unsigned long long cap_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // launched speculatively right behind the size pass, before the host knows the total: if the files
    // do not fit the buffer kept from earlier calls the whole grid backs off and the host launches again
    if (EMIT && offsets[gridDim.x] > cap_bytes) return;
    SizeSmem& ss = *reinterpret_cast<SizeSmem*>(smem_raw);
    EmitSmem& sm = *reinterpret_cast<EmitSmem*>(smem_raw);
    uint8_t* sm_kind = EMIT ? sm.kind : ss.kind;
    uint32_t* warp_tmp = EMIT ? sm.warp_tmp : ss.warp_tmp;
    const int tid = threadIdx.x;
    const uint4 job = jobs[blockIdx.x];
    const uint4* src = reinterpret_cast<const uint4*>(pool_voxels + (size_t)job.z * VOXELS_IN_AREA);
    uint4* area_kinds = reinterpret_cast<uint4*>(kinds + (size_t)blockIdx.x * ROWS);
    if (!EMIT) {
        for (int i = tid; i < ROWS; i += 256) ss.rows[i] = src[i];
        __syncthreads();
        // 1. kind of every row, lanes on consecutive rows (conflict-free 16-byte reads)
        for (int r = tid; r < ROWS; r += 256) {
            const uint4 v = ss.rows[r];
            uint8_t c = LIT;
            if (r > 0 && r < ROWS - 1) {
                if (same_row(v, ss.rows[r - 1])) c = M16;
                else if (r >= 16 && same_row(v, ss.rows[r - 16])) c = M256;
            }
            ss.kind[r] = c;
        }
        __syncthreads();
        if (tid < ROWS / 16) area_kinds[tid] = reinterpret_cast<const uint4*>(ss.kind)[tid];
    } else {
        if (tid < ROWS / 16) reinterpret_cast<uint4*>(sm.kind)[tid] = area_kinds[tid];
        __syncthreads();
    }
    // everything below works on the kinds only, thread t owning rows 8t .. 8t+7
    const int r0 = tid * 8;
    uint8_t kind[8];
    {
        const uint2 k8 = *reinterpret_cast<const uint2*>(&sm_kind[r0]);
#pragma unroll
        for (int k = 0; k < 8; ++k) kind[k] = (uint8_t)((k < 4 ? k8.x >> (8 * k) : k8.y >> (8 * (k - 4))) & 0xFFu);
    }
    const uint8_t kind_before = r0 > 0 ? sm_kind[r0 - 1] : (uint8_t)0xFF;
    const uint8_t kind_after = r0 + 8 < ROWS ? sm_kind[r0 + 8] : (uint8_t)0xFF;

    // 2. start row of the segment (maximal run of one kind) every row belongs to
    uint32_t last_start = 0;  // row + 1 of the last boundary among my rows, 0 = none
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if ((k == 0 ? kind_before : kind[k - 1]) != kind[k]) last_start = (uint32_t)(r0 + k) + 1u;
    uint32_t run_start = block_carry_max(last_start, warp_tmp, tid);
    uint16_t start[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if ((k == 0 ? kind_before : kind[k - 1]) != kind[k]) run_start = (uint32_t)(r0 + k) + 1u;
        start[k] = (uint16_t)(run_start - 1u);
    }
    if (EMIT) {
        uint4 packed;
        packed.x = start[0] | ((uint32_t)start[1] << 16);
        packed.y = start[2] | ((uint32_t)start[3] << 16);
        packed.z = start[4] | ((uint32_t)start[5] << 16);
        packed.w = start[6] | ((uint32_t)start[7] << 16);
        *reinterpret_cast<uint4*>(&sm.seg_start[r0]) = packed;
    }

    // 3. bytes every segment contributes, accounted to its END row.  A literal segment opens a
    //    sequence (token + length extension + literals); a match segment adds offset + extension,
    //    and its own token if no literals precede it.  Row 0 carries the 3 bincode header bytes.
    uint32_t contrib[8];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int r = r0 + k;
        contrib[k] = 0;
        if ((k == 7 ? kind_after : kind[k + 1]) != kind[k]) {  // segment end (0xFF after the last row)
            const uint32_t s = start[k], len_rows = (uint32_t)r - s + 1u;
            if (kind[k] == LIT) {
                const uint32_t L = 16u * len_rows + (s == 0 ? 3u : 0u);
                contrib[k] = 1u + ext_bytes(L) + L;
            } else {
                const bool own_token = sm_kind[s - 1] != LIT;  // s >= 1: row 0 is always literal
                contrib[k] = (own_token ? 1u : 0u) + 2u + ext_bytes(16u * len_rows - 4u);
            }
        }
        mine += contrib[k];
    }
    uint32_t block_bytes;
    uint32_t pos = block_carry_sum(mine, warp_tmp, tid, &block_bytes);
    if (!EMIT) {
        if (tid == 0) sizes[blockIdx.x] = 4u + block_bytes;  // compress_prepend_size: u32 LE size first
        return;
    }

    uint8_t* file = out + offsets[blockIdx.x];
    uint8_t* block = file + 4;
    if (tid == 0) {
        file[0] = (uint8_t)(PAYLOAD & 0xFF);
        file[1] = (uint8_t)((PAYLOAD >> 8) & 0xFF);
        file[2] = (uint8_t)((PAYLOAD >> 16) & 0xFF);
        file[3] = (uint8_t)(PAYLOAD >> 24);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int r = r0 + k;
        if (contrib[k]) {
            sm.seg_pos[r] = pos;
            if (kind[k] == LIT) {
                const uint32_t s = start[k];
                const uint32_t L = 16u * ((uint32_t)r - s + 1u) + (s == 0 ? 3u : 0u);
                sm.lit_base[s] = pos + 1u + ext_bytes(L);
            }
        }
        pos += contrib[k];
    }
    __syncthreads();

    // 4. tokens, extensions, offsets: one thread per sequence (the match segment's end row, or the
    //    last row for the closing literals-only sequence)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int r = r0 + k;
        if (!contrib[k]) continue;
        if (kind[k] != LIT) {
            const uint32_t s = start[k], ml = 16u * ((uint32_t)r - s + 1u) - 4u;
            const uint32_t offset = kind[k] == M16 ? 16u : 256u;
            uint8_t* p;
            if (sm.kind[s - 1] == LIT) {  // literals, then this match
                const uint32_t ls = sm.seg_start[s - 1];
                const uint32_t L = 16u * (s - ls) + (ls == 0 ? 3u : 0u);
                uint8_t* t = block + sm.seg_pos[s - 1];
                *t = (uint8_t)((min(L, 15u) << 4) | min(ml, 15u));
                if (L >= 15u) put_ext(t + 1, L);
                p = block + sm.seg_pos[r];
            } else {
                p = block + sm.seg_pos[r];
                *p++ = (uint8_t)min(ml, 15u);
            }
            *p++ = (uint8_t)(offset & 0xFF);
            *p++ = (uint8_t)(offset >> 8);
            if (ml >= 15u) put_ext(p, ml);
        } else if (r == ROWS - 1) {  // closing sequence: literals only
            const uint32_t s = start[k];
            const uint32_t L = 16u * ((uint32_t)r - s + 1u) + (s == 0 ? 3u : 0u);
            uint8_t* t = block + sm.seg_pos[r];
            *t = (uint8_t)(min(L, 15u) << 4);
            if (L >= 15u) put_ext(t + 1, L);
        }
    }
    // 5. literal rows, lanes on consecutive rows
    for (int r = tid; r < ROWS; r += 256) {
        if (sm.kind[r] != LIT) continue;
        const uint32_t s = sm.seg_start[r];
        uint8_t* p = block + sm.lit_base[s] + 16u * ((uint32_t)r - s) + (s == 0 ? 3u : 0u);
        if (r == 0) {  // bincode: slice length 32768 as a varint (marker 251 + u16 LE)
            p[-3] = 251;
            p[-2] = (uint8_t)(VOXELS_IN_AREA & 0xFF);
            p[-1] = (uint8_t)(VOXELS_IN_AREA >> 8);
        }
        const uint4 v = src[r];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if ((reinterpret_cast<uintptr_t>(p) & 3u) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<uint32_t*>(p)[q] = w[q];
        } else {
#pragma unroll
            for (int b = 0; b < 16; ++b) p[b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
        }
    }
}

// exclusive scan of the file sizes into 64-bit offsets; offsets[n] = total
__global__ void __launch_bounds__(1024) codec_offsets_kernel(const uint32_t* __restrict__ sizes, int n,
                                                             unsigned long long* __restrict__ offsets) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const unsigned long long c = i < n ? sizes[i] : 0ull;
        unsigned long long inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned long long w = s_w[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, w, d);
                if (lane >= d) w += o;
            }
            s_w[lane] = w;
        }
        __syncthreads();
        const unsigned long long before = carry + (wid ? s_w[wid - 1] : 0ull) + inc - c;
        if (i < n) offsets[i] = before;
        __syncthreads();
        if (tid == 1023) carry = before + c;
        __syncthreads();
    }
    if (tid == 0) offsets[n] = carry;
}

// ------------------------------------------------------------------------------------- decode
constexpr uint32_t STAGE_BYTES = 4096;  // window of the compressed file in shared memory (most files fit whole)
constexpr uint32_t WIN = 512;           // bytes of the most recent output kept in shared memory

// One warp per file, expanding STRAIGHT INTO THE AREA'S POOL SLOT.  Shared memory holds the staged
// compressed file and a 512-byte window of the most recent output, not the 32 KiB area: an SM keeps
// 32 files in flight instead of 6 -- the walk over the LZ4 sequences is a chain of dependent steps
// per file, and files in flight are what hides it.  Payload byte p (bincode: FB 00 80, then 32 768
// variant bytes) lives in hdr[p] for p < 3 and in vox[p - 3] after that, so voxel rows are 16-byte
// aligned in the slot; the window is indexed by (p - 3) mod 512, so they are aligned there too.
//   * a match whose offset fits the window (every match of the row-granular files this library
//     writes: offsets 16 and 256) is ONE parallel pass: output byte j is window byte
//     from + (j mod offset), whatever the length -- no rounds, nothing is read back;
//   * a longer offset (any valid LZ4 file is accepted) reads back what this warp has written, with
//     ld.global.cg behind a __syncwarp, in doubling rounds;
//   * voxel ids are checked where they enter the stream (literals); matches only repeat checked bytes.
struct DecodeSmem {
    uint8_t in[STAGE_BYTES];
    alignas(16) uint8_t win[WIN];
    uint8_t hdr[16];
};

__device__ __forceinline__ uint32_t win_slot(uint32_t p) { return (p + WIN - 3u) & (WIN - 1u); }

__device__ __forceinline__ uint32_t payload_get(const uint8_t* hdr, const uint8_t* vox, uint32_t p) {
    return p < 3u ? hdr[p] : __ldcg(vox + (p - 3u));
}
__device__ __forceinline__ void payload_put(uint8_t* hdr, uint8_t* vox, uint32_t p, uint32_t v) {
    if (p < 3u) hdr[p] = (uint8_t)v;
    else vox[p - 3u] = (uint8_t)v;
}

// payload[dst .. dst + n) = payload[src .. src + n), src + n <= dst, read back from the slot.
__device__ __forceinline__ void match_copy_far(uint8_t* hdr, uint8_t* vox, uint32_t dst, uint32_t src, uint32_t n, int lane) {
    uint32_t done = 0;
    if (src >= 3u && (((dst - 3u) | (src - 3u)) & 15u) == 0u) {
        const uint32_t n16 = n >> 4;
        uint4* d = reinterpret_cast<uint4*>(vox + (dst - 3u));
        const uint4* f = reinterpret_cast<const uint4*>(vox + (src - 3u));
        for (uint32_t k = lane; k < n16; k += 32) d[k] = __ldcg(f + k);
        done = n16 << 4;
    }
    for (uint32_t k = done + lane; k < n; k += 32) payload_put(hdr, vox, dst + k, payload_get(hdr, vox, src + k));
}

// LZ4 length extension: bytes are added to `acc` up to and including the first one that is not 255.
// The warp reads 32 bytes at a time (a 12 KiB match has 48 of them): one ballot finds the terminator,
// one reduction adds what precedes it.  false = the stream ends first, or the length is absurd.
__device__ __forceinline__ bool length_extension(const uint8_t* src, uint32_t n, uint32_t& ip, uint32_t& acc, int lane) {
    while (true) {
        const uint32_t pos = ip + (uint32_t)lane;
        const uint32_t b = pos < n ? src[pos] : 0u;  // past the end: reads as a terminator, caught below
        const unsigned stop = __ballot_sync(0xFFFFFFFFu, b != 255u);
        if (stop) {
            const int first = __ffs(stop) - 1;
            acc += __reduce_add_sync(0xFFFFFFFFu, lane <= first ? b : 0u);
            ip += (uint32_t)first + 1u;
            return ip <= n && acc <= PAYLOAD + 255u;
        }
        acc += 32u * 255u;
        ip += 32u;
        if (acc > PAYLOAD + 255u) return false;
    }
}

// status[i] = 1 if the file expands to a well-formed AreaDTO (then the voxels are in the area's pool
// slot), 0 where the reference logs an error and generates the area instead
// (world_persistence.rs:66-70): bad size prefix, malformed LZ4 stream, wrong slice length, voxel id
// >= 27 -- the slot then holds garbage and the host releases it.
__global__ void __launch_bounds__(32) codec_decode_kernel(const uint4* __restrict__ jobs,
                                                          const uint8_t* __restrict__ files,
                                                          const unsigned long long* __restrict__ offsets,
                                                          uint8_t* __restrict__ pool_voxels,
                                                          uint8_t* __restrict__ status) {
    __shared__ __align__(16) DecodeSmem sm;
    const int lane = threadIdx.x;
    const uint4 job = jobs[blockIdx.x];
    const uint8_t* file = files + offsets[blockIdx.x];
    const unsigned long long n64 = offsets[blockIdx.x + 1] - offsets[blockIdx.x];
    bool ok = n64 >= 4 && n64 <= 0x7FFFFFFFull;
    const uint32_t n = ok ? (uint32_t)n64 : 0u;
    if (ok) {  // decompress_size_prepended: u32 LE uncompressed size
        const uint32_t size = file[0] | (file[1] << 8) | (file[2] << 16) | ((uint32_t)file[3] << 24);
        ok = size == PAYLOAD;
    }
    // Files are packed back to back, so a file starts at any byte: whole 16-byte pieces of the
    // underlying buffer are staged and the parse reads at the file's position inside them.  A file
    // longer than the stage is parsed through a sliding window that is refilled between sequences
    // (the bytes a sequence header can span are bounded); its literals are copied from global memory.
    const uint32_t head = (uint32_t)((uintptr_t)file & 15u);
    const uint4* f16 = reinterpret_cast<const uint4*>(file - head);
    const uint32_t pieces = (head + n + 15u) >> 4;
    const uint8_t* src = sm.in + head;  // src[t] = file byte t while t < wend
    uint32_t wend = 0;                  // file bytes [.., wend) are staged
    auto restage = [&](uint32_t pos) {
        const uint32_t first = (head + pos) >> 4;  // piece that holds file byte pos
        const uint32_t count = min(pieces - first, STAGE_BYTES / 16u);
        __syncwarp();
        for (uint32_t i = lane; i < count; i += 32) reinterpret_cast<uint4*>(sm.in)[i] = __ldg(f16 + first + i);
        __syncwarp();
        src = sm.in + head - 16u * first;  // may point before sm.in: only [pos, wend) is ever read
        wend = min(n, 16u * (first + count) - head);
    };
    constexpr uint32_t HEADER_SPAN = 320;  // token + 2 x 129 length bytes + offset, with room to spare
    if (ok) restage(0);
    if (lane < 4) sm.hdr[lane] = 0;
    __syncwarp();
    uint8_t* vox = pool_voxels + (size_t)job.z * VOXELS_IN_AREA;
    uint32_t ip = 4, op = 0;
    bool bad_voxel = false;   // a literal byte that is no Voxel variant (this lane's share)
    bool recheck = false;     // a match copied header bytes into the voxels: check them all at the end
    while (ok && ip < n) {
        if (wend < n && ip + HEADER_SPAN > wend) restage(ip);
        const uint32_t token = src[ip++];
        uint32_t lit = token >> 4;
        if (lit == 15u && !length_extension(src, n, ip, lit, lane)) { ok = false; break; }
        if (lit > n - ip || lit > PAYLOAD - op) { ok = false; break; }
        const uint8_t* lsrc = ip + lit <= wend ? src : file;  // long literal runs of a long file: from global memory
        for (uint32_t k = lane; k < lit; k += 32) {
            const uint32_t v = lsrc[ip + k], p = op + k;
            payload_put(sm.hdr, vox, p, v);
            bad_voxel = bad_voxel || (p >= 3u && v >= (uint32_t)V_VARIANTS);
            if (lit - k <= WIN) sm.win[win_slot(p)] = (uint8_t)v;
        }
        ip += lit;
        op += lit;
        if (ip == n) break;  // last sequence: literals only
        if (n - ip < 2u) { ok = false; break; }
        if (wend < n && ip + HEADER_SPAN > wend) restage(ip);
        const uint32_t offset = src[ip] | (src[ip + 1] << 8);
        ip += 2;
        if (offset == 0u || offset > op) { ok = false; break; }
        uint32_t len = (token & 15u) + 4u;
        if ((token & 15u) == 15u && !length_extension(src, n, ip, len, lane)) { ok = false; break; }
        if (len > PAYLOAD - op) { ok = false; break; }
        __syncwarp();  // the literals (and everything before them) are in place
        const uint32_t from = op - offset, end = op + len;
        recheck = recheck || from < 3u;
        if (offset <= WIN) {
            // output byte j = payload[from + j mod offset], all of which are in the window
            const bool rows = (offset & 15u) == 0u && op >= 3u && ((op - 3u) & 15u) == 0u;
            const bool pow2 = (offset & (offset - 1u)) == 0u;  // 16 and 256, the offsets K8 writes
            const uint32_t omask = offset - 1u;
            auto wrap = [&](uint32_t j) { return pow2 ? (j & omask) : j % offset; };
            uint32_t done = 0;
            if (rows) {
                const uint32_t n16 = len >> 4;
                uint4* d = reinterpret_cast<uint4*>(vox + (op - 3u));
                for (uint32_t k = lane; k < n16; k += 32)
                    d[k] = *reinterpret_cast<const uint4*>(sm.win + win_slot(from + wrap(16u * k)));
                done = n16 << 4;
            }
            for (uint32_t j = done + lane; j < len; j += 32)
                payload_put(sm.hdr, vox, op + j, sm.win[win_slot(from + wrap(j))]);
            // the window moves on to [end - WIN, end): every lane works out its 16 bytes from the old
            // window, then all write.  Lanes whose 16 bytes lie before the match keep what is there.
            const uint32_t p0 = end - WIN + 16u * (uint32_t)lane;  // may wrap below 0: only p >= op is written
            const int32_t rel = (int32_t)(p0 - op);               // where the block starts inside the match
            const bool touched = rel + 16 > 0;
            const bool whole = rel >= 0 && rows && (len & 15u) == 0u;  // row-aligned and inside
            uint4 fresh = make_uint4(0, 0, 0, 0);
            if (touched) {
                if (whole) {
                    fresh = *reinterpret_cast<const uint4*>(sm.win + win_slot(from + wrap((uint32_t)rel)));
                } else {
                    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int b = 0; b < 16; ++b)
                        if (rel + b >= 0)
                            w[b >> 2] |= (uint32_t)sm.win[win_slot(from + wrap((uint32_t)(rel + b)))] << (8 * (b & 3));
                    fresh = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            __syncwarp();
            if (touched) {
                if (whole) {
                    *reinterpret_cast<uint4*>(sm.win + win_slot(p0)) = fresh;
                } else {
                    const uint32_t w[4] = {fresh.x, fresh.y, fresh.z, fresh.w};
#pragma unroll
                    for (int b = 0; b < 16; ++b)
                        if (rel + b >= 0) sm.win[win_slot(p0 + (uint32_t)b)] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
                }
            }
        } else {
            // An overlapping match repeats the last `offset` bytes.  Copy in rounds: each round copies
            // from the start of the pattern as much as is already final (offset, then twice that, ...),
            // so every round is a plain non-overlapping copy.
            uint32_t done = 0;
            while (done < len) {
                const uint32_t chunk = min(len - done, offset + done);
                match_copy_far(sm.hdr, vox, op + done, from, chunk, lane);
                done += chunk;
                __syncwarp();
            }
            const uint32_t tail = min(len, WIN);  // the window follows
            for (uint32_t k = lane; k < tail; k += 32) {
                const uint32_t p = end - tail + k;
                sm.win[win_slot(p)] = (uint8_t)payload_get(sm.hdr, vox, p);
            }
        }
        __syncwarp();
        op = end;
    }
    __syncwarp();
    ok = ok && op == PAYLOAD;
    // bincode: Box<[Voxel]> = varint length (FB 00 80) + one variant byte per voxel, each < 27
    ok = ok && sm.hdr[0] == 251 && sm.hdr[1] == (VOXELS_IN_AREA & 0xFF) && sm.hdr[2] == (VOXELS_IN_AREA >> 8);
    bool valid = !bad_voxel;
    if (ok && recheck) {
        const uint4* voxels = reinterpret_cast<const uint4*>(vox);
        const uint32_t limit = 0x01010101u * (uint32_t)(V_VARIANTS - 1);
        for (uint32_t w = lane; w < VOXELS_IN_AREA / 16; w += 32) {
            const uint4 v = __ldcg(voxels + w);
            // any byte > 26 is not a Voxel variant (per-byte unsigned compare)
            valid = valid && (__vcmpgtu4(v.x, limit) | __vcmpgtu4(v.y, limit) | __vcmpgtu4(v.z, limit) |
                              __vcmpgtu4(v.w, limit)) == 0u;
        }
    }
    ok = ok && __all_sync(0xFFFFFFFFu, valid);
    if (lane == 0) status[blockIdx.x] = ok ? 1 : 0;
}

}  // namespace

void launch_codec_sizes(const uint4* jobs, int n, const uint8_t* pool_voxels, uint8_t* kinds, uint32_t* sizes,
                        unsigned long long* offsets, cudaStream_t stream) {
    if (n <= 0) return;
    cudaFuncSetAttribute(codec_encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SizeSmem));
    codec_encode_kernel<false><<<n, 256, sizeof(SizeSmem), stream>>>(jobs, pool_voxels, kinds, sizes, nullptr, nullptr, 0ull);
    codec_offsets_kernel<<<1, 1024, 0, stream>>>(sizes, n, offsets);
}

void launch_codec_emit(const uint4* jobs, int n, const uint8_t* pool_voxels, uint8_t* kinds,
                       const unsigned long long* offsets, uint8_t* out, unsigned long long cap_bytes, cudaStream_t stream) {
    if (n <= 0) return;
    codec_encode_kernel<true><<<n, 256, sizeof(EmitSmem), stream>>>(jobs, pool_voxels, kinds, nullptr, offsets, out, cap_bytes);
}

void launch_codec_decode(const uint4* jobs, int n, const uint8_t* files, const unsigned long long* offsets,
                         uint8_t* pool_voxels, uint8_t* status, cudaStream_t stream) {
    if (n <= 0) return;
    codec_decode_kernel<<<n, 32, 0, stream>>>(jobs, files, offsets, pool_voxels, status);
}

}  // namespace vx
