// Oracle tooling only: eager stand-in for boost/numeric/ublas/vector_proxy.hpp. See storage.hpp.
#ifndef ORACLE_SHIM_UBLAS_VECTOR_PROXY_HPP
#define ORACLE_SHIM_UBLAS_VECTOR_PROXY_HPP
#include <boost/numeric/ublas/vector.hpp>
namespace boost { namespace numeric { namespace ublas {

template<class V, class I> class vector_indirect_proxy : public vector_expression< vector_indirect_proxy<V,I> > {
    V* _v; I _idx;
  public:
    typedef typename V::value_type value_type;
    typedef std::size_t size_type;
    vector_indirect_proxy(V& v, const I& idx) : _v(&v), _idx(idx) { }
    size_type size() const { return _idx.size(); }
    const value_type& operator()(size_type i) const { return (*_v)(_idx(i)); }
    const value_type& operator[](size_type i) const { return (*_v)(_idx(i)); }
    value_type& operator()(size_type i) { return const_cast<value_type&>((*_v)(_idx(i))); }
    value_type& operator[](size_type i) { return const_cast<value_type&>((*_v)(_idx(i))); }
    template<class E> vector_indirect_proxy& operator=(const vector_expression<E>& ve) {
        vector<value_type> tmp(ve); for(size_type i=0; i!=size(); ++i) { (*this)(i)=tmp(i); } return *this; }
    vector_indirect_proxy& operator=(const vector_indirect_proxy& o) {
        vector<value_type> tmp(o); for(size_type i=0; i!=size(); ++i) { (*this)(i)=tmp(i); } return *this; }
};
template<class V> using vector_range = vector_indirect_proxy<V,range>;
template<class V> using vector_slice = vector_indirect_proxy<V,slice>;

template<class T> vector_range< vector<T> > project(vector<T>& v, const range& r) { return vector_range< vector<T> >(v,r); }
template<class T> vector_range< const vector<T> > project(const vector<T>& v, const range& r) { return vector_range< const vector<T> >(v,r); }
template<class T> vector_slice< vector<T> > project(vector<T>& v, const slice& s) { return vector_slice< vector<T> >(v,s); }
template<class T> vector_slice< const vector<T> > project(const vector<T>& v, const slice& s) { return vector_slice< const vector<T> >(v,s); }

} } }
#endif
