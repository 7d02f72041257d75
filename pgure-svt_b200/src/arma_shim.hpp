// Minimal Armadillo-compatible containers so that PGURESVT<T1,T2>() keeps the reference's signature
// (src/pguresvt.hpp:17-38 takes arma::Cube / arma::Mat) without depending on Armadillo, which is not available
// in this image.  Column-major, owning or borrowing (the bridge of the reference wraps numpy memory without a
// copy: _pguresvt.pyx:72-93 -> Cube(ptr, r, c, s, copy_aux_mem=false, strict=false)).
// If real Armadillo is present, define PGURESVT_USE_ARMADILLO before including pguresvt.hpp and this header is skipped.
#ifndef PGURESVT_ARMA_SHIM_HPP
#define PGURESVT_ARMA_SHIM_HPP
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace arma
{
typedef unsigned long long uword;

template <typename T>
class Mat
{
public:
    uword n_rows = 0, n_cols = 0, n_elem = 0;
    Mat() {}
    Mat(uword r, uword c) { set_size(r, c); }
    Mat(T *aux, uword r, uword c, bool copy_aux_mem = true, bool /*strict*/ = false)
    {
        if (copy_aux_mem)
        {
            set_size(r, c);
            std::copy(aux, aux + n_elem, own.begin());
        }
        else
        {
            n_rows = r, n_cols = c, n_elem = r * c, ext = aux;
        }
    }
    void set_size(uword r, uword c)
    {
        n_rows = r, n_cols = c, n_elem = r * c;
        own.assign(n_elem, T());
        ext = nullptr;
    }
    void zeros() { std::fill(memptr(), memptr() + n_elem, T()); }
    T *memptr() { return ext ? ext : own.data(); }
    const T *memptr() const { return ext ? ext : own.data(); }
    T &operator()(uword r, uword c) { return memptr()[r + n_rows * c]; }
    const T &operator()(uword r, uword c) const { return memptr()[r + n_rows * c]; }

private:
    std::vector<T> own;
    T *ext = nullptr;
};

template <typename T>
class Cube
{
public:
    uword n_rows = 0, n_cols = 0, n_slices = 0, n_elem = 0;
    Cube() {}
    Cube(uword r, uword c, uword s) { set_size(r, c, s); }
    Cube(T *aux, uword r, uword c, uword s, bool copy_aux_mem = true, bool /*strict*/ = false)
    {
        if (copy_aux_mem)
        {
            set_size(r, c, s);
            std::copy(aux, aux + n_elem, own.begin());
        }
        else
        {
            n_rows = r, n_cols = c, n_slices = s, n_elem = r * c * s, ext = aux;
        }
    }
    void set_size(uword r, uword c, uword s)
    {
        n_rows = r, n_cols = c, n_slices = s, n_elem = r * c * s;
        own.assign(n_elem, T());
        ext = nullptr;
    }
    void zeros() { std::fill(memptr(), memptr() + n_elem, T()); }
    T *memptr() { return ext ? ext : own.data(); }
    const T *memptr() const { return ext ? ext : own.data(); }
    T *slice_memptr(uword s) { return memptr() + n_rows * n_cols * s; }
    const T *slice_memptr(uword s) const { return memptr() + n_rows * n_cols * s; }
    T &operator()(uword r, uword c, uword s) { return memptr()[r + n_rows * (c + n_cols * s)]; }
    const T &operator()(uword r, uword c, uword s) const { return memptr()[r + n_rows * (c + n_cols * s)]; }
    T min() const { return *std::min_element(memptr(), memptr() + n_elem); }
    T max() const { return *std::max_element(memptr(), memptr() + n_elem); }

private:
    std::vector<T> own;
    T *ext = nullptr;
};

typedef Cube<double> cube;
typedef Mat<double> mat;
} // namespace arma
#endif
