// Host-side entropy parser: Mobiclip frame bytes -> packed per-macroblock arrays (include/mobicuda.h).
// This is the half of DecodeVXS2 (MobiclipDecoder.cs:97-259, "MD") that stays on the CPU: every branch
// that consumes bits depends only on bits and parser state, never on pixels.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/mobicuda.h"

namespace mobi {

// Append-only array whose capacity is topped up once per macroblock (room(k) guarantees k more unchecked appends),
// so the per-coefficient / per-leaf hot loops carry no capacity checks.
template <class T>
struct Arr {
    std::vector<T> v;
    size_t n = 0;
    void clear() { n = 0; }
    size_t size() const { return n; }
    void room(size_t k) { if (n + k > v.size()) v.resize((n + k) * 2); }
    void push_back(const T& x) { v[n++] = x; }
    T& operator[](size_t i) { return v[i]; }
    const T* data() const { return v.data(); }
};

struct ParsedFrame {
    mobi_frame_hdr hdr;
    Arr<mobi_mb> mbs;
    Arr<mobi_part> parts;
    Arr<mobi_op> ops;
    Arr<mobi_coef> coefs;
    Arr<uint32_t> intra;
    void clear() { mbs.clear(); parts.clear(); ops.clear(); coefs.clear(); intra.clear(); }
    // worst case of one macroblock: 64 leaves (2x2), 27 intra ops, 6 x 64 coefficients
    void room_for_mb() { mbs.room(1); parts.room(64); ops.room(32); coefs.room(384); intra.room(1); }
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
    // The ring is owned by the caller (the batch): it tells the parser how many pictures it holds before every parse, so
    // that the two counts cannot drift apart (a reset of the ring, a step that failed after its parse succeeded).
    void set_pictures(int n) { st_.decoded = n < 0 ? 0 : n > 6 ? 6 : n; }
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
        bool qtab_clean = true;   // every matrix index in qtab is in range (always, unless ModsDS runs with q < 12)
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
