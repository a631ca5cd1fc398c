// Build shim for the reference arm: Boost is not installed in this image and the single-node
// build of the reference needs exactly one Boost macro.  BOOST_STRONG_TYPEDEF(T, D) declares a
// distinct type D that wraps a T (explicit construction from T, implicit conversion back).
#pragma once
#define BOOST_STRONG_TYPEDEF(T, D)                                              \
  struct D {                                                                    \
    T t;                                                                        \
    explicit D(const T& v) noexcept : t(v) {}                                   \
    D() noexcept : t() {}                                                       \
    D(const D&) = default;                                                      \
    D& operator=(const D&) = default;                                           \
    D& operator=(const T& v) noexcept { t = v; return *this; }                  \
    operator const T&() const { return t; }                                     \
    operator T&() { return t; }                                                 \
    bool operator==(const D& o) const { return t == o.t; }                      \
    bool operator<(const D& o) const { return t < o.t; }                        \
  };
