// Host side of the B200 FFV1 decoder: reads the ConfigurationRecord (once per stream) with a byte-wise binary range
// decoder. Field order and checks follow parameters::Parse, /root/reference/Source/Lib/CoDec/FFV1/FFV1_Parameters.cpp:23-183;
// the tables are rebuilt as parameters::QuantizationTable does (:222-253); the range decoder is rangecoder::b / ::u / ::s
// (FFV1_RangeCoder.cpp:25-171) including its late renormalisation (the byte is fetched when the next bin needs it).
#include <cstring>

#include "../../include/b200enc.h"
#include "ffv1_dec.h"
#include "ffv1_host.h"

namespace b200 {
namespace {

class HostRangeDecoder {
  public:
    HostRangeDecoder(const uint8_t* p, size_t n, const uint8_t* one) : beg_(p), cur_(p), end_(p + n) {
        if (n) low_ = *cur_;
        range_ = 0xFF;
        cur_++;
        std::memcpy(one_, one, 256);
        zero_[0] = 0;
        for (int i = 1; i < 256; i++) zero_[i] = (uint8_t)(256 - one_[256 - i]);
    }
    bool bin(uint8_t& st) {
        if (range_ < 0x100) {
            low_ <<= 8;
            if (cur_ > end_) { underrun_ = true; return false; }
            if (cur_ < end_) low_ |= *cur_;
            range_ <<= 8;
            cur_++;
        }
        uint32_t r1 = (range_ * st) >> 8;
        range_ -= r1;
        if (low_ < range_) { st = zero_[st]; return false; }
        low_ -= range_;
        range_ = r1;
        st = one_[st];
        return true;
    }
    uint32_t u(uint8_t* st) {
        if (bin(st[0])) return 0;
        int e = 0;
        while (bin(st[1 + (e < 9 ? e : 9)])) {
            if (++e > 31) { underrun_ = true; return 0; }
        }
        uint32_t a = 1;
        for (int i = e - 1; i >= 0; i--) a = (a << 1) | (bin(st[22 + (i < 9 ? i : 9)]) ? 1u : 0u);
        return a;
    }
    int32_t s(uint8_t* st) {
        if (bin(st[0])) return 0;
        int e = 0;
        while (bin(st[1 + (e < 9 ? e : 9)])) {
            if (++e > 31) { underrun_ = true; return 0; }
        }
        int32_t a = 1;
        for (int i = e - 1; i >= 0; i--) a = (a << 1) | (bin(st[22 + (i < 9 ? i : 9)]) ? 1 : 0);
        return bin(st[11 + (e < 10 ? e : 10)]) ? -a : a;
    }
    bool underrun() const { return underrun_; }

  private:
    const uint8_t *beg_, *cur_, *end_;
    uint32_t low_ = 0, range_ = 0xFF;
    bool underrun_ = false;
    uint8_t one_[256], zero_[256];
};

int bad(std::string* err, const char* what) {
    if (err) *err = what;
    return B200_ERR_INVALID;
}

}  // namespace

int parse_config_record(const uint8_t* rec, size_t n, Ffv1DecStream* S, std::string* err) {
    if (!rec || n < 5) return bad(err, "FFV1-HEADER-END:1");
    S->crc_ok = crc32_mpeg(rec, n) == 0;
    if (!S->crc_ok) return bad(err, "FFV1-HEADER-configuration_record_crc_parity:1");
    uint8_t def[256];
    default_one_state(def);
    HostRangeDecoder E(rec, n - 4, def);
    uint8_t st[32];
    std::memset(st, 128, sizeof st);
    S->version = (int)E.u(st);
    if (S->version <= 1) return bad(err, "FFV1-HEADER-version-OUTOFBAND:1");
    if (S->version != 3) return bad(err, S->version == 2 ? "FFV1-HEADER-version-EXPERIMENTAL:1" : "FFV1-HEADER-version-LATERVERSION:1");
    S->micro = (int)E.u(st);
    if (S->micro < 4) return bad(err, "FFV1-HEADER-micro_version-EXPERIMENTAL:1");
    S->coder_type_sent = (int)E.u(st);
    if (S->coder_type_sent > 2) return bad(err, "FFV1-HEADER-coder_type:1");
    std::memcpy(S->one_state, def, 256);
    if (S->coder_type_sent == 2) {
        for (int i = 1; i < 256; i++) {
            int v = (int)def[i] + E.s(st);
            if (v < 0 || v > 255) return bad(err, "FFV1-HEADER-state_transition_delta:1");
            S->one_state[i] = (uint8_t)v;
        }
    }
    S->zero_state[0] = 0;
    for (int i = 1; i < 256; i++) S->zero_state[i] = (uint8_t)(256 - S->one_state[256 - i]);
    S->colorspace = (int)E.u(st);
    if (S->colorspace > 1) return bad(err, "FFV1-HEADER-colorspace_type:1");
    S->bits = (int)E.u(st);
    if (S->bits > 64) return bad(err, "FFV1-HEADER-bits_per_raw_sample:1");
    if (S->bits == 0) S->bits = 8;
    S->chroma_planes = E.bin(st[0]);
    S->log2_h = (int)E.u(st);
    S->log2_v = (int)E.u(st);
    S->alpha = E.bin(st[0]);
    S->num_h = (int)E.u(st) + 1;
    S->num_v = (int)E.u(st) + 1;
    S->nsets = (int)E.u(st);
    if (S->nsets > 8 || S->nsets < 1) return bad(err, "FFV1-HEADER-quant_table_count:1");
    S->qtab.assign((size_t)S->nsets * 5 * 256, 0);
    for (int i = 0; i < S->nsets; i++) {
        int scale = 1;
        for (int j = 0; j < 5; j++) {
            int16_t* t = S->qtab.data() + ((size_t)i * 5 + j) * 256;
            uint8_t qs[32];
            std::memset(qs, 128, sizeof qs);
            int v = 0;
            for (int k = 0; k < 128;) {
                uint32_t len_minus1 = E.u(qs);
                if (k + len_minus1 >= 128 || E.underrun()) return bad(err, "FFV1-HEADER-QuantizationTable-len:1");
                for (uint32_t a = 0; a <= len_minus1; a++) t[k++] = (int16_t)(scale * v);
                v++;
            }
            for (int k = 1; k < 128; k++) t[256 - k] = (int16_t)-t[k];
            t[128] = (int16_t)-t[127];
            scale *= 2 * v - 1;
            if (scale > 32768) return bad(err, "FFV1-HEADER-QuantizationTable-scale:1");
        }
        S->nctx[i] = (scale + 1) >> 1;
    }
    for (int i = 0; i < S->nsets; i++) {
        if (E.bin(st[0])) {
            S->states_coded = true;
            return bad(err, "coded initial states are not supported by the B200 decoder");
        }
    }
    S->ec = (int)E.u(st);
    if (S->ec > 1) return bad(err, "FFV1-HEADER-ec:1");
    S->intra = (int)E.u(st);
    if (S->intra > 1) return bad(err, "FFV1-HEADER-intra:1");
    if (E.underrun()) return bad(err, "FFV1-HEADER-END:1");
    return 0;
}

}  // namespace b200
