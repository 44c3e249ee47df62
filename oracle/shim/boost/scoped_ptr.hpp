// Oracle tooling only: stand-in for boost/scoped_ptr.hpp.
#ifndef ORACLE_SHIM_BOOST_SCOPED_PTR_HPP
#define ORACLE_SHIM_BOOST_SCOPED_PTR_HPP
#include <memory>
namespace boost {
template<class T> class scoped_ptr {
    T* _p;
    scoped_ptr(const scoped_ptr&);
    scoped_ptr& operator=(const scoped_ptr&);
  public:
    explicit scoped_ptr(T* p=0) : _p(p) { }
    ~scoped_ptr() { delete _p; }
    void reset(T* p=0) { if(p!=_p) { delete _p; _p=p; } }
    T& operator*() const { return *_p; }
    T* operator->() const { return _p; }
    T* get() const { return _p; }
    explicit operator bool() const { return _p!=0; }
    void swap(scoped_ptr& o) { T* t=_p; _p=o._p; o._p=t; }
};
}
#endif
