// TEST INFRASTRUCTURE -- not part of the product, never linked into libmobicuda.so.
//
// C#-semantics prelude for oracle/_ref: the handful of .NET types the reference decoder
// (LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs) touches, re-stated in C++ so that the
// syntax-level transliteration produced by oracle/build_ref.py compiles and keeps the
// reference's run-time behaviour:
//   * Arr<T>    = managed T[]: reference semantics, zero-initialised, bounds-checked
//                 (an out-of-range index throws, as IndexOutOfRangeException would, so the
//                 reference's catch-all at MobiclipDecoder.cs:325 aborts the frame the same way)
//   * Bitmap/BitmapData/Color = the 32bpp ARGB pixel container used at MobiclipDecoder.cs:260-323
// Nothing here is decoder logic.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include <initializer_list>
#include <cmath>
#include <functional>
#include <string>
#include <utility>

typedef uint8_t byte;
typedef int8_t sbyte;
typedef uint16_t ushort;
typedef uint32_t uint;
typedef uint64_t ulong;

struct Exception { };
struct NotImplementedException : Exception { };
struct IndexOutOfRangeException : Exception { };
struct NullReferenceException : Exception { };
struct ArgumentException : Exception { };
struct KeyNotFoundException : Exception { };

template <class T>
struct Arr {
    std::shared_ptr<std::vector<T>> p;
    int Length = 0;
    Arr() {}
    Arr(std::nullptr_t) {}
    Arr(std::initializer_list<T> il) : p(std::make_shared<std::vector<T>>(il)), Length((int)il.size()) {}
    static Arr New(long long n) {
        Arr a;
        if (n < 0) throw Exception();
        a.p = std::make_shared<std::vector<T>>((size_t)n);
        a.Length = (int)n;
        return a;
    }
    T& operator[](long long i) const {
        if (!p) throw NullReferenceException();
        if (i < 0 || i >= (long long)Length) throw IndexOutOfRangeException();
        return (*p)[(size_t)i];
    }
    bool operator==(const Arr& o) const { return p == o.p; }
    bool operator!=(const Arr& o) const { return p != o.p; }
    T* raw() const { return p ? p->data() : nullptr; }
};

struct Array {
    template <class T>
    static void Copy(const Arr<T>& src, long long si, const Arr<T>& dst, long long di, long long n) {
        if (!src.p || !dst.p) throw NullReferenceException();
        if (n < 0 || si < 0 || di < 0 || si + n > src.Length || di + n > dst.Length) throw IndexOutOfRangeException();
        if (n) std::memmove(dst.raw() + di, src.raw() + si, (size_t)n * sizeof(T));
    }
};

// ---- System.Collections.Generic.List<T>, as far as BitWriter.cs uses it (reference semantics like the managed type) ----
template <class T>
struct List {
    std::shared_ptr<std::vector<T>> p = std::make_shared<std::vector<T>>();
    void Add(T x) { p->push_back(x); }
    void AddRange(const List& o) { p->insert(p->end(), o.p->begin(), o.p->end()); }
    void AddRange(const Arr<T>& a) { if (!a.p) throw NullReferenceException(); p->insert(p->end(), a.p->begin(), a.p->end()); }
    void Clear() { p->clear(); }
    Arr<T> ToArray() const {
        Arr<T> a = Arr<T>::New((long long)p->size());
        if (!p->empty()) std::memcpy(a.raw(), p->data(), p->size() * sizeof(T));
        return a;
    }
};

// ---- System.Drawing stand-ins (pixel container only) ----
struct Rectangle { int X, Y, Width, Height; Rectangle(int x, int y, int w, int h) : X(x), Y(y), Width(w), Height(h) {} };
enum class ImageLockMode { WriteOnly };
enum class PixelFormat { Format32bppArgb };
struct BitmapData { byte* Scan0; int Stride; };
struct Color {
    int a, r, g, b;
    static Color FromArgb(int r, int g, int b) {
        if ((unsigned)r > 255u || (unsigned)g > 255u || (unsigned)b > 255u) throw Exception();  // ArgumentException in .NET
        return Color{255, r, g, b};
    }
    int ToArgb() const { return (int)(((uint)a << 24) | ((uint)r << 16) | ((uint)g << 8) | (uint)b); }
};
struct Bitmap {
    std::shared_ptr<std::vector<byte>> px;
    int Width = 0, Height = 0;
    Bitmap() {}
    Bitmap(std::nullptr_t) {}
    Bitmap(int w, int h) : px(std::make_shared<std::vector<byte>>((size_t)w * h * 4)), Width(w), Height(h) {}
    BitmapData LockBits(Rectangle, ImageLockMode, PixelFormat) { return BitmapData{px->data(), Width * 4}; }
    void UnlockBits(BitmapData) {}
    bool IsNull() const { return !px; }
};
// ---- what the container classes (LibMobiclip/Containers) touch: a seekable byte stream, Dictionary, ASCII strings, Math ----
// System.IO.Stream over memory (MemoryStream semantics: Read returns what is left, 0 at the end; Position may sit past the end;
// writes grow the stream).
struct CsStream {
    std::vector<byte> buf;
    long long Position = 0;
    CsStream() {}
    CsStream(const byte* d, size_t n) : buf(d, d + n) {}
    long long Length() const { return (long long)buf.size(); }
    int Read(const Arr<byte>& dst, long long off, long long count) {
        if (!dst.p) throw NullReferenceException();
        if (off < 0 || count < 0 || off + count > dst.Length) throw IndexOutOfRangeException();
        long long left = (long long)buf.size() - Position;
        if (left < 0) left = 0;
        const long long n = count < left ? count : left;
        if (n > 0) std::memcpy(dst.raw() + off, buf.data() + Position, (size_t)n);
        Position += n;
        return (int)n;
    }
    int ReadByte() { if (Position >= (long long)buf.size() || Position < 0) return -1; return buf[(size_t)Position++]; }
    void Write(const Arr<byte>& src, long long off, long long count) {
        if (!src.p) throw NullReferenceException();
        if (off < 0 || count < 0 || off + count > src.Length) throw IndexOutOfRangeException();
        if (Position + count > (long long)buf.size()) buf.resize((size_t)(Position + count));
        if (count) std::memcpy(buf.data() + Position, src.raw() + off, (size_t)count);
        Position += count;
    }
    void WriteByte(byte b) { if (Position + 1 > (long long)buf.size()) buf.resize((size_t)(Position + 1)); buf[(size_t)Position++] = b; }
    void Flush() {}
};
// Dictionary<K, V>: enumeration in insertion order (what .NET does while nothing has been removed), Add throws on a duplicate key
template <class K, class V>
struct Dictionary {
    std::vector<std::pair<K, V>> kv;
    int find(const K& k) const { for (size_t i = 0; i < kv.size(); i++) if (kv[i].first == k) return (int)i; return -1; }
    void Add(const K& k, const V& v) { if (find(k) >= 0) throw ArgumentException(); kv.emplace_back(k, v); }
    bool ContainsKey(const K& k) const { return find(k) >= 0; }
    V& operator[](const K& k) { const int i = find(k); if (i < 0) throw KeyNotFoundException(); return kv[(size_t)i].second; }
    void Clear() { kv.clear(); }
    std::vector<V> Values() const { std::vector<V> v; for (auto& e : kv) v.push_back(e.second); return v; }
};
typedef std::string CsString;
struct Encoding {
    struct Ascii { CsString GetString(const Arr<byte>& d, int off, int n) const { CsString s; for (int i = 0; i < n; i++) s.push_back((char)(d[off + i] & 0x7F)); return s; } };
    static constexpr Ascii ASCII{};
};
struct Math {
    static double Floor(double x) { return std::floor(x); }
    static double Log(double a, double newBase) { return std::log(a) / std::log(newBase); }
};
#define null nullptr
