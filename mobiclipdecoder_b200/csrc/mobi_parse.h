// Host-side entropy parser: Mobiclip frame bytes -> packed per-macroblock arrays (include/mobicuda.h).
// This is the half of DecodeVXS2 (MobiclipDecoder.cs:97-259, "MD") that stays on the CPU: every branch
// that consumes bits depends only on bits and parser state, never on pixels.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/mobicuda.h"

namespace mobi {

struct ParsedFrame {
    mobi_frame_hdr hdr;
    std::vector<mobi_mb> mbs;
    std::vector<mobi_part> parts;
    std::vector<mobi_op> ops;
    std::vector<mobi_coef> coefs;
    std::vector<uint32_t> intra;
    void clear() { mbs.clear(); parts.clear(); ops.clear(); coefs.clear(); intra.clear(); }
    mobi_packed_frame view() const { return mobi_packed_frame{&hdr, mbs.data(), parts.data(), ops.data(), coefs.data(), intra.data()}; }
};

inline int stride_for(uint32_t w) { return w <= 256 ? 256 : w <= 512 ? 512 : 1024; }  // MD:50-52

class Parser {
public:
    Parser(uint32_t w, uint32_t h, int version);
    // Returns a mobi_status.  On success *offset advances like MobiclipDecoder.Offset and the
    // "pictures in the ring" count grows by one; on failure all state is restored.
    int parse(const uint8_t* data, int len, int* offset, ParsedFrame& out);
    void reset();                      // back to a freshly constructed decoder (no pictures, Quantizer 0)
    uint32_t quantizer() const { return st_.quant; }
    uint32_t yuv_format() const { return st_.yuvfmt; }
    int pictures() const { return st_.decoded; }
    const std::string& error() const { return err_; }
    uint32_t width() const { return W_; }
    uint32_t height() const { return H_; }
    int stride() const { return S_; }
    int version() const { return ver_; }

private:
    struct State {
        uint32_t quant = 0, yuvfmt = 0;
        uint32_t qtab[80] = {0};  // Internal[10..89]
        uint8_t ctx[40] = {0};    // Internal bytes 0..39: intra-mode context grid
        int decoded = 0;          // pictures currently in the ring (saturates at 6)
    };
    struct Bits;
    friend struct FrameParse;
    uint32_t W_, H_;
    int ver_, S_, mbw_, mbh_;
    State st_;
    std::vector<int32_t> mvc_;
    std::string err_;
};

}  // namespace mobi
