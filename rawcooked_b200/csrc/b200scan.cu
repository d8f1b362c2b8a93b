// Analysis-pass helpers on the B200 (include/b200scan.h): MD5 of whole files and the DPX padding-bit test.
//
//   MD5            RFC 1321 as /root/reference/Source/Lib/ThirdParty/md5/md5.c computes it for input_base::Hash
//                  (Source/Lib/Utils/FileIO/Input_Base.cpp:54-81). A message is one serial chain of 64-step blocks, so the
//                  parallelism is messages: k_md5 gives every file of the batch one lane.
//   padding bits   the test of dpx::ParseBuffer, Source/Lib/Uncompressed/DPX/DPX.cpp:500-608, restated per flavor:
//                  Filled A: one byte per sample word carries the padding bits (mask 0x3 for 10 bit, 0xF for 12 bit; which
//                  byte depends on the endianness); Packed: the last 32-bit word of every row, read big endian, masked with
//                  0xFFFFFFFF << (used bits % 32) when the row does not end on a word boundary.
#include "../../include/b200scan.h"

#include <cuda_runtime.h>

#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "ffv1_host.h"
#include "pixel_layouts.cuh"

void b200_set_error(const std::string& msg);

namespace {

int sfail(int code, const std::string& msg) { b200_set_error(msg); return code; }
int sfail_cuda(cudaError_t e, const char* what) {
    b200_set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? B200_ERR_NO_DEVICE : B200_ERR_CUDA;
}
#define SCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return sfail_cuda(e_, #x); } while (0)

// ------------------------------------------------------------------------------------------------------------------
// MD5
__device__ __forceinline__ uint32_t rotl(uint32_t v, int s) { return __funnelshift_l(v, v, s); }

#define MD5_F(x, y, z) ((z) ^ ((x) & ((y) ^ (z))))
#define MD5_G(x, y, z) ((y) ^ ((z) & ((x) ^ (y))))
#define MD5_H(x, y, z) ((x) ^ (y) ^ (z))
#define MD5_I(x, y, z) ((y) ^ ((x) | ~(z)))
#define MD5_STEP(f, a, b, c, d, x, t, s) (a) += f((b), (c), (d)) + (x) + (t); (a) = rotl((a), (s)); (a) += (b);

__device__ __forceinline__ void md5_block(uint32_t st[4], const uint32_t w[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
    MD5_STEP(MD5_F, a, b, c, d, w[0], 0xd76aa478, 7)  MD5_STEP(MD5_F, d, a, b, c, w[1], 0xe8c7b756, 12)
    MD5_STEP(MD5_F, c, d, a, b, w[2], 0x242070db, 17) MD5_STEP(MD5_F, b, c, d, a, w[3], 0xc1bdceee, 22)
    MD5_STEP(MD5_F, a, b, c, d, w[4], 0xf57c0faf, 7)  MD5_STEP(MD5_F, d, a, b, c, w[5], 0x4787c62a, 12)
    MD5_STEP(MD5_F, c, d, a, b, w[6], 0xa8304613, 17) MD5_STEP(MD5_F, b, c, d, a, w[7], 0xfd469501, 22)
    MD5_STEP(MD5_F, a, b, c, d, w[8], 0x698098d8, 7)  MD5_STEP(MD5_F, d, a, b, c, w[9], 0x8b44f7af, 12)
    MD5_STEP(MD5_F, c, d, a, b, w[10], 0xffff5bb1, 17) MD5_STEP(MD5_F, b, c, d, a, w[11], 0x895cd7be, 22)
    MD5_STEP(MD5_F, a, b, c, d, w[12], 0x6b901122, 7) MD5_STEP(MD5_F, d, a, b, c, w[13], 0xfd987193, 12)
    MD5_STEP(MD5_F, c, d, a, b, w[14], 0xa679438e, 17) MD5_STEP(MD5_F, b, c, d, a, w[15], 0x49b40821, 22)
    MD5_STEP(MD5_G, a, b, c, d, w[1], 0xf61e2562, 5)  MD5_STEP(MD5_G, d, a, b, c, w[6], 0xc040b340, 9)
    MD5_STEP(MD5_G, c, d, a, b, w[11], 0x265e5a51, 14) MD5_STEP(MD5_G, b, c, d, a, w[0], 0xe9b6c7aa, 20)
    MD5_STEP(MD5_G, a, b, c, d, w[5], 0xd62f105d, 5)  MD5_STEP(MD5_G, d, a, b, c, w[10], 0x02441453, 9)
    MD5_STEP(MD5_G, c, d, a, b, w[15], 0xd8a1e681, 14) MD5_STEP(MD5_G, b, c, d, a, w[4], 0xe7d3fbc8, 20)
    MD5_STEP(MD5_G, a, b, c, d, w[9], 0x21e1cde6, 5)  MD5_STEP(MD5_G, d, a, b, c, w[14], 0xc33707d6, 9)
    MD5_STEP(MD5_G, c, d, a, b, w[3], 0xf4d50d87, 14) MD5_STEP(MD5_G, b, c, d, a, w[8], 0x455a14ed, 20)
    MD5_STEP(MD5_G, a, b, c, d, w[13], 0xa9e3e905, 5) MD5_STEP(MD5_G, d, a, b, c, w[2], 0xfcefa3f8, 9)
    MD5_STEP(MD5_G, c, d, a, b, w[7], 0x676f02d9, 14) MD5_STEP(MD5_G, b, c, d, a, w[12], 0x8d2a4c8a, 20)
    MD5_STEP(MD5_H, a, b, c, d, w[5], 0xfffa3942, 4)  MD5_STEP(MD5_H, d, a, b, c, w[8], 0x8771f681, 11)
    MD5_STEP(MD5_H, c, d, a, b, w[11], 0x6d9d6122, 16) MD5_STEP(MD5_H, b, c, d, a, w[14], 0xfde5380c, 23)
    MD5_STEP(MD5_H, a, b, c, d, w[1], 0xa4beea44, 4)  MD5_STEP(MD5_H, d, a, b, c, w[4], 0x4bdecfa9, 11)
    MD5_STEP(MD5_H, c, d, a, b, w[7], 0xf6bb4b60, 16) MD5_STEP(MD5_H, b, c, d, a, w[10], 0xbebfbc70, 23)
    MD5_STEP(MD5_H, a, b, c, d, w[13], 0x289b7ec6, 4) MD5_STEP(MD5_H, d, a, b, c, w[0], 0xeaa127fa, 11)
    MD5_STEP(MD5_H, c, d, a, b, w[3], 0xd4ef3085, 16) MD5_STEP(MD5_H, b, c, d, a, w[6], 0x04881d05, 23)
    MD5_STEP(MD5_H, a, b, c, d, w[9], 0xd9d4d039, 4)  MD5_STEP(MD5_H, d, a, b, c, w[12], 0xe6db99e5, 11)
    MD5_STEP(MD5_H, c, d, a, b, w[15], 0x1fa27cf8, 16) MD5_STEP(MD5_H, b, c, d, a, w[2], 0xc4ac5665, 23)
    MD5_STEP(MD5_I, a, b, c, d, w[0], 0xf4292244, 6)  MD5_STEP(MD5_I, d, a, b, c, w[7], 0x432aff97, 10)
    MD5_STEP(MD5_I, c, d, a, b, w[14], 0xab9423a7, 15) MD5_STEP(MD5_I, b, c, d, a, w[5], 0xfc93a039, 21)
    MD5_STEP(MD5_I, a, b, c, d, w[12], 0x655b59c3, 6) MD5_STEP(MD5_I, d, a, b, c, w[3], 0x8f0ccc92, 10)
    MD5_STEP(MD5_I, c, d, a, b, w[10], 0xffeff47d, 15) MD5_STEP(MD5_I, b, c, d, a, w[1], 0x85845dd1, 21)
    MD5_STEP(MD5_I, a, b, c, d, w[8], 0x6fa87e4f, 6)  MD5_STEP(MD5_I, d, a, b, c, w[15], 0xfe2ce6e0, 10)
    MD5_STEP(MD5_I, c, d, a, b, w[6], 0xa3014314, 15) MD5_STEP(MD5_I, b, c, d, a, w[13], 0x4e0811a1, 21)
    MD5_STEP(MD5_I, a, b, c, d, w[4], 0xf7537e82, 6)  MD5_STEP(MD5_I, d, a, b, c, w[11], 0xbd3af235, 10)
    MD5_STEP(MD5_I, c, d, a, b, w[2], 0x2ad7d2bb, 15) MD5_STEP(MD5_I, b, c, d, a, w[9], 0xeb86d391, 21)
    st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}

// byte `off` of the padded message: data, then 0x80, then zeros (the length words are put in by the caller)
__device__ __forceinline__ uint32_t md5_tail_byte(const uint8_t* p, uint64_t len, uint64_t off) {
    return off < len ? p[off] : (off == len ? 0x80u : 0u);
}

// one lane per message: lane i hashes [base + off[i], + len[i])
__global__ void __launch_bounds__(32) k_md5(const uint8_t* __restrict__ base, const uint64_t* __restrict__ off,
                                            const uint64_t* __restrict__ len, int n, uint32_t* __restrict__ digests) {
    const int i = blockIdx.x * 32 + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = base + off[i];
    const uint64_t L = len[i];
    uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    const uint64_t nfull = L >> 6;
    uint32_t w[16];
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        // the next block is asked for before the current one is hashed: the load latency hides behind the 64 steps
        uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0, n2 = n0, n3 = n0;
        if (nfull) { n0 = __ldg(q); n1 = __ldg(q + 1); n2 = __ldg(q + 2); n3 = __ldg(q + 3); }
        for (uint64_t b = 0; b < nfull; b++) {
            w[0] = n0.x; w[1] = n0.y; w[2] = n0.z; w[3] = n0.w; w[4] = n1.x; w[5] = n1.y; w[6] = n1.z; w[7] = n1.w;
            w[8] = n2.x; w[9] = n2.y; w[10] = n2.z; w[11] = n2.w; w[12] = n3.x; w[13] = n3.y; w[14] = n3.z; w[15] = n3.w;
            if (b + 1 < nfull) {
                const uint4* r = q + (b + 1) * 4;
                n0 = __ldg(r); n1 = __ldg(r + 1); n2 = __ldg(r + 2); n3 = __ldg(r + 3);
            }
            md5_block(st, w);
        }
    } else {
        for (uint64_t b = 0; b < nfull; b++) {
            const uint8_t* r = p + b * 64;
#pragma unroll
            for (int k = 0; k < 16; k++) w[k] = (uint32_t)r[4 * k] | ((uint32_t)r[4 * k + 1] << 8) | ((uint32_t)r[4 * k + 2] << 16) | ((uint32_t)r[4 * k + 3] << 24);
            md5_block(st, w);
        }
    }
    // the rest, 0x80, zeros, and the length in bits in the last eight bytes of the last block
    const uint32_t rem = (uint32_t)(L & 63);
    const int ntail = rem < 56 ? 1 : 2;
    for (int t = 0; t < ntail; t++) {
        const uint64_t o = nfull * 64 + (uint64_t)t * 64;
#pragma unroll
        for (int k = 0; k < 16; k++)
            w[k] = md5_tail_byte(p, L, o + 4 * k) | (md5_tail_byte(p, L, o + 4 * k + 1) << 8) | (md5_tail_byte(p, L, o + 4 * k + 2) << 16) |
                   (md5_tail_byte(p, L, o + 4 * k + 3) << 24);
        if (t == ntail - 1) { w[14] = (uint32_t)(L << 3); w[15] = (uint32_t)(L >> 29); }
        md5_block(st, w);
    }
    digests[4 * i + 0] = st[0]; digests[4 * i + 1] = st[1]; digests[4 * i + 2] = st[2]; digests[4 * i + 3] = st[3];
}

// ------------------------------------------------------------------------------------------------------------------
// padding bits
struct PadArgs {
    const uint8_t* in;
    uint8_t* out;                        // masked copy or null
    unsigned long long* count;           // [n]
    unsigned long long* first;           // [n], preset to all ones
    size_t frame_bytes;
    uint32_t row_bytes, eol_off, eol_mask, rows;
    uint32_t umask;                      // Filled: mask on a unit loaded little endian
    int n;
};

// Filled A: every sample word (32 bit for 10-bit RGB, 16 bit for a 12-bit component) has its padding bits in one byte.
// A streaming pass: 16-byte loads (and stores, when the masked copy is wanted) over the 16-byte aligned body of the
// payload, the few units before and after it one by one. HBM-bound: one read of the payload (+ one write).
template <class Unit>
__device__ __forceinline__ void pad_unit(const PadArgs& A, const Unit* p, Unit* o, size_t u, unsigned long long& cnt, unsigned long long& first, uint32_t tb) {
    const Unit v = (Unit)(p[u] & (Unit)A.umask);
    if (o) o[u] = v;
    if (v) {
        cnt++;
        const unsigned long long at = (unsigned long long)u * sizeof(Unit) + tb;
        first = at < first ? at : first;
    }
}
template <class Unit>
__global__ void __launch_bounds__(256) k_padding_filled(const PadArgs A) {
    const int f = blockIdx.y;
    const uint8_t* base = A.in + (size_t)f * A.frame_bytes;
    const Unit* p = reinterpret_cast<const Unit*>(base);
    Unit* o = A.out ? reinterpret_cast<Unit*>(A.out + (size_t)f * A.frame_bytes) : nullptr;
    uint32_t tb = 0;                                             // offset of the tested byte inside the unit
    for (uint32_t m = A.umask; !(m & 0xFFu); m >>= 8) tb++;
    const uint32_t m32 = sizeof(Unit) == 4 ? A.umask : (A.umask | (A.umask << 16));
    const size_t head = (size_t)((16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15);      // bytes before the aligned body
    const size_t nvec = A.frame_bytes > head ? (A.frame_bytes - head) / 16 : 0;
    // the masked copy is vector-stored only when it is as aligned as the source
    const bool vec_out = !o || ((reinterpret_cast<uintptr_t>(A.out + (size_t)f * A.frame_bytes) & 15) == (reinterpret_cast<uintptr_t>(base) & 15));
    unsigned long long cnt = 0, first = ~0ull;
    const size_t gtid = (size_t)blockIdx.x * 256 + threadIdx.x, gstride = (size_t)gridDim.x * 256;
    if (vec_out) {
        const uint4* pv = reinterpret_cast<const uint4*>(base + head);
        uint4* ov = o ? reinterpret_cast<uint4*>(A.out + (size_t)f * A.frame_bytes + head) : nullptr;
        for (size_t k = gtid; k < nvec; k += gstride) {
            uint4 v = __ldg(pv + k);
            v.x &= m32; v.y &= m32; v.z &= m32; v.w &= m32;
            if (ov) ov[k] = v;
            if (v.x | v.y | v.z | v.w) {
                const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 3; j >= 0; j--) {
                    if (!wd[j]) continue;
                    const unsigned long long at = head + k * 16 + 4 * j;
                    if (sizeof(Unit) == 4) { cnt++; first = at + tb < first ? at + tb : first; }
                    else {
                        if (wd[j] >> 16) { cnt++; first = at + 2 + tb < first ? at + 2 + tb : first; }
                        if (wd[j] & 0xFFFFu) { cnt++; first = at + tb < first ? at + tb : first; }
                    }
                }
            }
        }
        // units before and after the body
        const size_t uhead = head / sizeof(Unit), utail0 = (head + nvec * 16) / sizeof(Unit), units = A.frame_bytes / sizeof(Unit);
        if (blockIdx.x == 0) {
            for (size_t u = threadIdx.x; u < uhead && u < units; u += 256) pad_unit<Unit>(A, p, o, u, cnt, first, tb);
            for (size_t u = utail0 + threadIdx.x; u < units; u += 256) pad_unit<Unit>(A, p, o, u, cnt, first, tb);
        }
    } else {
        const size_t units = A.frame_bytes / sizeof(Unit);
        for (size_t u = gtid; u < units; u += gstride) pad_unit<Unit>(A, p, o, u, cnt, first, tb);
    }
    for (int s = 16; s; s >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, first, s);
        first = t < first ? t : first;
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(&A.count[f], cnt);
        atomicMin(&A.first[f], first);
    }
}

// Packed: the last 32-bit word of every row, read big endian (DPX.cpp:507-521, :575-583)
__global__ void __launch_bounds__(256) k_padding_packed(const PadArgs A) {
    const size_t r = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= (size_t)A.n * A.rows) return;
    const int f = (int)(r / A.rows);
    const size_t o = (size_t)f * A.frame_bytes + (r % A.rows) * A.row_bytes + A.eol_off;
    const uint32_t wv = *reinterpret_cast<const uint32_t*>(A.in + o);
    const uint32_t v = __byte_perm(wv, 0, 0x0123) & A.eol_mask;
    if (A.out) *reinterpret_cast<uint32_t*>(A.out + o) = __byte_perm(v, 0, 0x0123);
    if (v) {
        atomicAdd(&A.count[f], 1ull);
        atomicMin(&A.first[f], (unsigned long long)(o - (size_t)f * A.frame_bytes));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The frame as FFmpeg's rawvideo encoder emits it for the flavor (what a `-f framemd5` output hashes,
// /root/reference/Source/CLI/Output.cpp:312-332): the pix_fmt of libavcodec's dpx / tiff decoder, planes back to back, rows
// without padding. kind 0: rgb24, 1: rgb48 little endian, 2: rgb48 big endian, 3: gbrp 16-bit little endian (planes G, B, R)
__global__ void __launch_bounds__(256) k_rawframe(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t frame_bytes,
                                                  uint32_t row_bytes, size_t raw_bytes, int w, int h, int layout, int kind) {
    const int f = blockIdx.y;
    const uint8_t* fin = in + (size_t)f * frame_bytes;
    uint8_t* fo = out + (size_t)f * raw_bytes;
    const size_t npx = (size_t)w * h;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < npx; i += (size_t)gridDim.x * 256) {
        const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
        int r, g, b;
        b200::load_rgb(fin + (size_t)y * row_bytes, layout, x, r, g, b);
        if (kind == 0) {
            uint8_t* q = fo + 3 * i;
            q[0] = (uint8_t)r; q[1] = (uint8_t)g; q[2] = (uint8_t)b;
        } else if (kind == 3) {
            uint16_t* q = reinterpret_cast<uint16_t*>(fo);
            q[i] = (uint16_t)g; q[npx + i] = (uint16_t)b; q[2 * npx + i] = (uint16_t)r;
        } else {
            uint16_t* q = reinterpret_cast<uint16_t*>(fo) + 3 * i;
            if (kind == 2) { r = (int)__byte_perm((uint32_t)r, 0, 0x4401); g = (int)__byte_perm((uint32_t)g, 0, 0x4401); b = (int)__byte_perm((uint32_t)b, 0, 0x4401); }
            q[0] = (uint16_t)r; q[1] = (uint16_t)g; q[2] = (uint16_t)b;
        }
    }
}

int rawvideo_kind(int layout) {
    switch (layout) {
        case B200_DPX_RGB_8: case B200_TIFF_RGB_8: return 0;
        case B200_DPX_RGB_16_LE: case B200_TIFF_RGB_16_LE: return 1;
        case B200_DPX_RGB_16_BE: case B200_TIFF_RGB_16_BE: return 2;
        case B200_DPX_RGB_10_FILLED_A_LE: case B200_DPX_RGB_10_FILLED_A_BE:
        case B200_DPX_RGB_12_FILLED_A_LE: case B200_DPX_RGB_12_FILLED_A_BE: case B200_DPX_RGB_12_PACKED_BE: return 3;
    }
    return -1;
}

}  // namespace

struct b200_scan {
    int device = 0;
    int max_items = 0;
    size_t max_bytes = 0;
    uint8_t* d_data = nullptr;
    uint8_t* d_out = nullptr;            // masked payloads of the host entry point (allocated on first use)
    uint8_t* d_raw = nullptr; size_t raw_cap = 0;   // rawvideo frames of the framemd5 entry points (allocated on first use)
    uint64_t *d_off = nullptr, *d_len = nullptr;
    uint32_t* d_dig = nullptr;
    unsigned long long *d_cnt = nullptr, *d_first = nullptr;
    uint64_t* h_meta = nullptr;          // pinned [off n][len n]
    uint32_t* h_dig = nullptr;           // pinned
    unsigned long long* h_res = nullptr; // pinned [cnt n][first n]
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    uint64_t stats[4] = {0};
};

extern "C" {

int b200_scan_open(int32_t device, int32_t max_items, size_t max_bytes, b200_scan** out) {
    if (!out || max_items < 1) return sfail(B200_ERR_INVALID, "bad argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t de = cudaGetDeviceCount(&ndev);
    if (de != cudaSuccess || ndev == 0) return sfail(B200_ERR_NO_DEVICE, "no CUDA device: the B200 analysis helpers have no CPU fallback");
    if (device < 0 || device >= ndev) return sfail(B200_ERR_INVALID, "bad device ordinal");
    SCU(cudaSetDevice(device));
    b200_scan* S = new (std::nothrow) b200_scan;
    if (!S) return sfail(B200_ERR_INVALID, "out of host memory");
    S->device = device; S->max_items = max_items; S->max_bytes = max_bytes;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    if (max_bytes) ok(cudaMalloc(reinterpret_cast<void**>(&S->d_data), max_bytes + (size_t)max_items * 16 + 64));
    ok(cudaMalloc(reinterpret_cast<void**>(&S->d_off), (size_t)max_items * 8));
    ok(cudaMalloc(reinterpret_cast<void**>(&S->d_len), (size_t)max_items * 8));
    ok(cudaMalloc(reinterpret_cast<void**>(&S->d_dig), (size_t)max_items * 16));
    ok(cudaMalloc(reinterpret_cast<void**>(&S->d_cnt), (size_t)max_items * 8));
    ok(cudaMalloc(reinterpret_cast<void**>(&S->d_first), (size_t)max_items * 8));
    ok(cudaMallocHost(reinterpret_cast<void**>(&S->h_meta), (size_t)max_items * 16));
    ok(cudaMallocHost(reinterpret_cast<void**>(&S->h_dig), (size_t)max_items * 16));
    ok(cudaMallocHost(reinterpret_cast<void**>(&S->h_res), (size_t)max_items * 16));
    ok(cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking));
    ok(cudaEventCreate(&S->ev[0]));
    ok(cudaEventCreate(&S->ev[1]));
    if (e != cudaSuccess) { b200_scan_close(S); return sfail_cuda(e, "scan allocation"); }
    *out = S;
    return 0;
}

void b200_scan_close(b200_scan* S) {
    if (!S) return;
    cudaSetDevice(S->device);
    if (S->stream) { cudaStreamSynchronize(S->stream); cudaStreamDestroy(S->stream); }
    for (auto v : S->ev) if (v) cudaEventDestroy(v);
    for (void* p : {(void*)S->d_data, (void*)S->d_out, (void*)S->d_raw, (void*)S->d_off, (void*)S->d_len, (void*)S->d_dig, (void*)S->d_cnt, (void*)S->d_first})
        if (p) cudaFree(p);
    for (void* p : {(void*)S->h_meta, (void*)S->h_dig, (void*)S->h_res}) if (p) cudaFreeHost(p);
    delete S;
}

int b200_md5_device(b200_scan* S, const void* d_base, const size_t* off, const size_t* len, int32_t n, uint8_t* digests, void* stream) {
    if (!S || !d_base || !off || !len || !digests) return sfail(B200_ERR_INVALID, "null argument");
    if (n < 1 || n > S->max_items) return sfail(B200_ERR_INVALID, "n out of range");
    SCU(cudaSetDevice(S->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint64_t total = 0;
    for (int i = 0; i < n; i++) { S->h_meta[i] = off[i]; S->h_meta[n + i] = len[i]; total += len[i]; }
    SCU(cudaMemcpyAsync(S->d_off, S->h_meta, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    SCU(cudaMemcpyAsync(S->d_len, S->h_meta + n, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    SCU(cudaEventRecord(S->ev[0], s));
    k_md5<<<(n + 31) / 32, 32, 0, s>>>(static_cast<const uint8_t*>(d_base), S->d_off, S->d_len, n, S->d_dig);
    SCU(cudaGetLastError());
    SCU(cudaEventRecord(S->ev[1], s));
    SCU(cudaMemcpyAsync(S->h_dig, S->d_dig, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
    SCU(cudaStreamSynchronize(s));
    std::memcpy(digests, S->h_dig, (size_t)n * 16);      // state words little endian = the digest bytes
    float ms = 0;
    cudaEventElapsedTime(&ms, S->ev[0], S->ev[1]);
    S->stats[0] = (uint64_t)(ms * 1000.0f); S->stats[1] = total;
    return 0;
}

int b200_md5_host(b200_scan* S, const uint8_t* const* data, const size_t* len, int32_t n, uint8_t* digests) {
    if (!S || !data || !len || !digests) return sfail(B200_ERR_INVALID, "null argument");
    if (n < 1 || n > S->max_items) return sfail(B200_ERR_INVALID, "n out of range");
    SCU(cudaSetDevice(S->device));
    std::vector<size_t> off(n);
    size_t total = 0;
    for (int i = 0; i < n; i++) { off[i] = total; total += (len[i] + 15) & ~(size_t)15; }
    if (total > S->max_bytes + (size_t)S->max_items * 16) return sfail(B200_ERR_OVERFLOW, "more bytes than the handle was opened for");
    for (int i = 0; i < n; i++)
        if (len[i]) SCU(cudaMemcpyAsync(S->d_data + off[i], data[i], len[i], cudaMemcpyHostToDevice, S->stream));
    return b200_md5_device(S, S->d_data, off.data(), len, n, digests, S->stream);
}

static int padding_plan(uint32_t width, uint32_t height, int32_t layout, PadArgs* A, bool* packed, bool* any) {
    std::memset(A, 0, sizeof *A);
    const int bits = b200::layout_bits(layout);
    if (!bits || layout >= 32) return sfail(B200_ERR_INVALID, "the padding test is defined for the DPX layouts (0..7)");
    A->row_bytes = (uint32_t)b200::layout_row_bytes(width, layout);
    A->frame_bytes = (size_t)A->row_bytes * height;
    A->rows = height;
    *any = true;
    switch (layout) {
        case B200_DPX_RGB_10_FILLED_A_LE: *packed = false; A->umask = 0x00000003u; break;      // byte 0 of the little-endian word
        case B200_DPX_RGB_10_FILLED_A_BE: *packed = false; A->umask = 0x03000000u; break;      // byte 3 (DPX.cpp:528-534: i += Step - 1)
        case B200_DPX_RGB_12_FILLED_A_LE: *packed = false; A->umask = 0x000Fu; break;
        case B200_DPX_RGB_12_FILLED_A_BE: *packed = false; A->umask = 0x0F00u; break;
        default: {                                                                             // Packed (8, 12 packed, 16 bit): DPX.cpp:507-521
            *packed = true;
            const uint64_t used = (uint64_t)width * bits * 3;
            const uint32_t rem = (uint32_t)(used % 32);
            if (!rem) { *any = false; break; }
            A->eol_off = (uint32_t)(used / 32) * 4;
            A->eol_mask = 0xFFFFFFFFu << rem;
            if (A->eol_off + 4 != A->row_bytes) return sfail(B200_ERR_INVALID, "row stride differs from the reference's (DPX.cpp:516-518)");
        }
    }
    return 0;
}

int b200_padding_device(b200_scan* S, uint32_t width, uint32_t height, int32_t layout, const void* d_payloads, int32_t n,
                        uint64_t* nonzero, uint64_t* first, void* d_masked, void* stream) {
    if (!S || !d_payloads || !nonzero || !first) return sfail(B200_ERR_INVALID, "null argument");
    if (n < 1 || n > S->max_items) return sfail(B200_ERR_INVALID, "n out of range");
    SCU(cudaSetDevice(S->device));
    PadArgs A;
    bool packed = false, any = true;
    if (int r = padding_plan(width, height, layout, &A, &packed, &any)) return r;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    A.in = static_cast<const uint8_t*>(d_payloads); A.out = static_cast<uint8_t*>(d_masked); A.count = S->d_cnt; A.first = S->d_first; A.n = n;
    SCU(cudaMemsetAsync(S->d_cnt, 0, (size_t)n * 8, s));
    SCU(cudaMemsetAsync(S->d_first, 0xFF, (size_t)n * 8, s));
    SCU(cudaEventRecord(S->ev[0], s));
    uint64_t read = 0;
    if (any) {
        if (packed) {
            if (A.out) SCU(cudaMemsetAsync(A.out, 0, (size_t)n * A.frame_bytes, s));
            const size_t rows = (size_t)n * A.rows;
            k_padding_packed<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(A);
            read = rows * 4;
        } else {
            const bool u32 = layout == B200_DPX_RGB_10_FILLED_A_LE || layout == B200_DPX_RGB_10_FILLED_A_BE;
            const size_t vecs = A.frame_bytes / 16 + 1;
            unsigned gx = (unsigned)((vecs + 256 * 4 - 1) / (256 * 4));
            const unsigned cap = (unsigned)((148 * 8 + n - 1) / n) > 8u ? (unsigned)((148 * 8 + n - 1) / n) : 8u;   // ~8 CTAs per SM over the batch
            if (gx > cap) gx = cap;
            if (gx < 1) gx = 1;
            if (u32) k_padding_filled<uint32_t><<<dim3(gx, n), 256, 0, s>>>(A);
            else k_padding_filled<uint16_t><<<dim3(gx, n), 256, 0, s>>>(A);
            read = (uint64_t)n * A.frame_bytes;
        }
        SCU(cudaGetLastError());
    } else if (A.out) {
        SCU(cudaMemsetAsync(A.out, 0, (size_t)n * A.frame_bytes, s));
    }
    SCU(cudaEventRecord(S->ev[1], s));
    SCU(cudaMemcpyAsync(S->h_res, S->d_cnt, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    SCU(cudaMemcpyAsync(S->h_res + n, S->d_first, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    SCU(cudaStreamSynchronize(s));
    for (int i = 0; i < n; i++) { nonzero[i] = S->h_res[i]; first[i] = S->h_res[n + i]; }
    float ms = 0;
    cudaEventElapsedTime(&ms, S->ev[0], S->ev[1]);
    S->stats[0] = (uint64_t)(ms * 1000.0f); S->stats[1] = read;
    return 0;
}

int b200_padding_host(b200_scan* S, uint32_t width, uint32_t height, int32_t layout, const uint8_t* const* payloads, int32_t n,
                      uint64_t* nonzero, uint64_t* first, uint8_t* const* masked) {
    if (!S || !payloads) return sfail(B200_ERR_INVALID, "null argument");
    if (n < 1 || n > S->max_items) return sfail(B200_ERR_INVALID, "n out of range");
    SCU(cudaSetDevice(S->device));
    const size_t fb = b200::layout_row_bytes(width, layout) * (size_t)height;
    if (!fb) return sfail(B200_ERR_INVALID, "unsupported layout");
    if (fb * n > S->max_bytes) return sfail(B200_ERR_OVERFLOW, "more bytes than the handle was opened for");
    bool want_out = false;
    if (masked) for (int i = 0; i < n; i++) want_out = want_out || masked[i];
    if (want_out && !S->d_out) SCU(cudaMalloc(reinterpret_cast<void**>(&S->d_out), S->max_bytes + 64));
    for (int i = 0; i < n; i++) SCU(cudaMemcpyAsync(S->d_data + (size_t)i * fb, payloads[i], fb, cudaMemcpyHostToDevice, S->stream));
    int r = b200_padding_device(S, width, height, layout, S->d_data, n, nonzero, first, want_out ? S->d_out : nullptr, S->stream);
    if (r) return r;
    if (want_out) {
        for (int i = 0; i < n; i++)
            if (masked[i]) SCU(cudaMemcpyAsync(masked[i], S->d_out + (size_t)i * fb, fb, cudaMemcpyDeviceToHost, S->stream));
        SCU(cudaStreamSynchronize(S->stream));
    }
    return 0;
}

size_t b200_rawvideo_bytes(uint32_t width, uint32_t height, int32_t layout) {
    const int k = rawvideo_kind(layout);
    return k < 0 ? 0 : (size_t)width * height * (k == 0 ? 3 : 6);
}

const char* b200_rawvideo_pix_fmt(int32_t layout) {
    switch (rawvideo_kind(layout)) {
        case 0: return "rgb24";
        case 1: return "rgb48le";
        case 2: return "rgb48be";
        case 3: return b200::layout_bits(layout) == 10 ? "gbrp10le" : "gbrp12le";
    }
    return "";
}

int b200_framemd5_device(b200_scan* S, uint32_t width, uint32_t height, int32_t layout, const void* d_payloads, int32_t n,
                         uint8_t* digests, void* stream) {
    if (!S || !d_payloads || !digests) return sfail(B200_ERR_INVALID, "null argument");
    if (n < 1 || n > S->max_items) return sfail(B200_ERR_INVALID, "n out of range");
    const int kind = rawvideo_kind(layout);
    if (kind < 0 || !width || !height) return sfail(B200_ERR_INVALID, "unsupported layout");
    SCU(cudaSetDevice(S->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t rb = (uint32_t)b200::layout_row_bytes(width, layout);
    const size_t fb = (size_t)rb * height, raw = b200_rawvideo_bytes(width, height, layout);
    std::vector<size_t> off(n), len(n, raw);
    // flavors whose payload already is the raw frame (tight rows of rgb24 / rgb48 in the file's byte order) are hashed in place
    const bool in_place = kind != 3 && fb == raw;
    if (in_place) {
        for (int i = 0; i < n; i++) off[i] = (size_t)i * fb;
        return b200_md5_device(S, d_payloads, off.data(), len.data(), n, digests, stream);
    }
    const size_t stride = (raw + 15) & ~(size_t)15;
    if (stride * n > S->raw_cap) {
        if (S->d_raw) cudaFree(S->d_raw);
        S->d_raw = nullptr; S->raw_cap = 0;
        SCU(cudaMalloc(reinterpret_cast<void**>(&S->d_raw), stride * (size_t)S->max_items + 64));
        S->raw_cap = stride * (size_t)S->max_items;
    }
    const size_t npx = (size_t)width * height;
    unsigned gx = (unsigned)((npx + 256 * 4 - 1) / (256 * 4));
    if (gx > 148 * 4) gx = 148 * 4;
    k_rawframe<<<dim3(gx, n), 256, 0, s>>>(static_cast<const uint8_t*>(d_payloads), S->d_raw, fb, rb, stride, (int)width, (int)height, layout, kind);
    SCU(cudaGetLastError());
    for (int i = 0; i < n; i++) off[i] = (size_t)i * stride;
    return b200_md5_device(S, S->d_raw, off.data(), len.data(), n, digests, stream);
}

int b200_framemd5_host(b200_scan* S, uint32_t width, uint32_t height, int32_t layout, const uint8_t* const* payloads, int32_t n,
                       uint8_t* digests) {
    if (!S || !payloads || !digests) return sfail(B200_ERR_INVALID, "null argument");
    if (n < 1 || n > S->max_items) return sfail(B200_ERR_INVALID, "n out of range");
    SCU(cudaSetDevice(S->device));
    const size_t fb = b200::layout_row_bytes(width, layout) * (size_t)height;
    if (!fb) return sfail(B200_ERR_INVALID, "unsupported layout");
    if (fb * n > S->max_bytes) return sfail(B200_ERR_OVERFLOW, "more bytes than the handle was opened for");
    for (int i = 0; i < n; i++) SCU(cudaMemcpyAsync(S->d_data + (size_t)i * fb, payloads[i], fb, cudaMemcpyHostToDevice, S->stream));
    return b200_framemd5_device(S, width, height, layout, S->d_data, n, digests, S->stream);
}

int b200_scan_stats(const b200_scan* S, uint64_t stats[4]) {
    if (!S || !stats) return sfail(B200_ERR_INVALID, "null argument");
    std::memcpy(stats, S->stats, sizeof S->stats);
    return 0;
}

}  // extern "C"
