// Oracle tooling only: stand-in for the boost::numeric::ublas subset used by the reference.
// Eager (no expression templates): every operator returns a concrete vector/matrix. Accumulation order of
// prod/inner_prod matches ublas's dense indexing loops (t=0; t+=a(i,k)*b(k) for k=0..n-1).
#ifndef ORACLE_SHIM_UBLAS_STORAGE_HPP
#define ORACLE_SHIM_UBLAS_STORAGE_HPP
#include <cstddef>
#include <cassert>
#include <algorithm>
#include <vector>
#include <iostream>
#include <type_traits>
#include <utility>
namespace boost { namespace numeric { namespace ublas {

template<class T> struct shim_void { typedef void type; };
// SFINAE-friendly: no promote_type member when X+Y is ill-formed, so operator templates drop out quietly.
template<class X, class Y, class Enable=void> struct promote_traits { };
template<class X, class Y> struct promote_traits<X,Y,
    typename shim_void<decltype(std::declval<X>()+std::declval<Y>())>::type> {
    typedef decltype(std::declval<X>()+std::declval<Y>()) promote_type; };

class range {
    std::size_t _start, _size;
  public:
    typedef std::size_t size_type;
    range() : _start(0), _size(0) { }
    range(size_type start, size_type stop) : _start(start), _size(stop-start) { }
    size_type start() const { return _start; }
    size_type size() const { return _size; }
    size_type operator()(size_type i) const { return _start+i; }
};
class slice {
    std::size_t _start; std::ptrdiff_t _stride; std::size_t _size;
  public:
    typedef std::size_t size_type;
    slice() : _start(0), _stride(0), _size(0) { }
    slice(size_type start, std::ptrdiff_t stride, size_type size) : _start(start), _stride(stride), _size(size) { }
    size_type start() const { return _start; }
    std::ptrdiff_t stride() const { return _stride; }
    size_type size() const { return _size; }
    size_type operator()(size_type i) const { return _start+i*_stride; }
};

template<class E> struct vector_expression {
    typedef E expression_type;
    const E& operator()() const { return *static_cast<const E*>(this); }
    E& operator()() { return *static_cast<E*>(this); }
};
template<class E> struct matrix_expression {
    typedef E expression_type;
    const E& operator()() const { return *static_cast<const E*>(this); }
    E& operator()() { return *static_cast<E*>(this); }
};

} } }
#endif
