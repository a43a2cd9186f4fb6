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

typedef uint8_t byte;
typedef int8_t sbyte;
typedef uint16_t ushort;
typedef uint32_t uint;
typedef uint64_t ulong;

struct Exception { };
struct NotImplementedException : Exception { };
struct IndexOutOfRangeException : Exception { };
struct NullReferenceException : Exception { };

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
#define null nullptr
