// Oracle tooling only: eager stand-in for boost/numeric/ublas/matrix_proxy.hpp. See storage.hpp.
#ifndef ORACLE_SHIM_UBLAS_MATRIX_PROXY_HPP
#define ORACLE_SHIM_UBLAS_MATRIX_PROXY_HPP
#include <boost/numeric/ublas/matrix.hpp>
#include <boost/numeric/ublas/vector_proxy.hpp>
namespace boost { namespace numeric { namespace ublas {

template<class M, class I1, class I2> class matrix_indirect_proxy : public matrix_expression< matrix_indirect_proxy<M,I1,I2> > {
    M* _m; I1 _i; I2 _j;
  public:
    typedef typename M::value_type value_type;
    typedef std::size_t size_type;
    matrix_indirect_proxy(M& m, const I1& i, const I2& j) : _m(&m), _i(i), _j(j) { }
    size_type size1() const { return _i.size(); }
    size_type size2() const { return _j.size(); }
    const value_type& operator()(size_type i, size_type j) const { return (*_m)(_i(i),_j(j)); }
    value_type& operator()(size_type i, size_type j) { return const_cast<value_type&>((*_m)(_i(i),_j(j))); }
    template<class E> matrix_indirect_proxy& operator=(const matrix_expression<E>& me) {
        matrix<value_type> tmp(me);
        for(size_type i=0; i!=size1(); ++i) { for(size_type j=0; j!=size2(); ++j) { (*this)(i,j)=tmp(i,j); } } return *this; }
    matrix_indirect_proxy& operator=(const matrix_indirect_proxy& o) {
        matrix<value_type> tmp(o);
        for(size_type i=0; i!=size1(); ++i) { for(size_type j=0; j!=size2(); ++j) { (*this)(i,j)=tmp(i,j); } } return *this; }
};
template<class M> using matrix_range = matrix_indirect_proxy<M,range,range>;
template<class M> using matrix_slice = matrix_indirect_proxy<M,slice,slice>;

template<class T> matrix_range< matrix<T> > project(matrix<T>& m, const range& r1, const range& r2) { return matrix_range< matrix<T> >(m,r1,r2); }
template<class T> matrix_range< const matrix<T> > project(const matrix<T>& m, const range& r1, const range& r2) { return matrix_range< const matrix<T> >(m,r1,r2); }
template<class T> matrix_slice< matrix<T> > project(matrix<T>& m, const slice& r1, const slice& r2) { return matrix_slice< matrix<T> >(m,r1,r2); }
template<class T> matrix_slice< const matrix<T> > project(const matrix<T>& m, const slice& r1, const slice& r2) { return matrix_slice< const matrix<T> >(m,r1,r2); }

template<class M> class matrix_row : public vector_expression< matrix_row<M> > {
    M* _m; std::size_t _i;
  public:
    typedef typename M::value_type value_type;
    typedef std::size_t size_type;
    matrix_row(M& m, size_type i) : _m(&m), _i(i) { }
    size_type size() const { return _m->size2(); }
    const value_type& operator()(size_type j) const { return (*_m)(_i,j); }
    const value_type& operator[](size_type j) const { return (*_m)(_i,j); }
    value_type& operator()(size_type j) { return const_cast<value_type&>((*_m)(_i,j)); }
    value_type& operator[](size_type j) { return const_cast<value_type&>((*_m)(_i,j)); }
    template<class E> matrix_row& operator=(const vector_expression<E>& ve) {
        vector<value_type> tmp(ve); for(size_type j=0; j!=size(); ++j) { (*this)(j)=tmp(j); } return *this; }
    matrix_row& operator=(const matrix_row& o) {
        vector<value_type> tmp(o); for(size_type j=0; j!=size(); ++j) { (*this)(j)=tmp(j); } return *this; }
};
template<class M> class matrix_column : public vector_expression< matrix_column<M> > {
    M* _m; std::size_t _j;
  public:
    typedef typename M::value_type value_type;
    typedef std::size_t size_type;
    matrix_column(M& m, size_type j) : _m(&m), _j(j) { }
    size_type size() const { return _m->size1(); }
    const value_type& operator()(size_type i) const { return (*_m)(i,_j); }
    const value_type& operator[](size_type i) const { return (*_m)(i,_j); }
    value_type& operator()(size_type i) { return const_cast<value_type&>((*_m)(i,_j)); }
    value_type& operator[](size_type i) { return const_cast<value_type&>((*_m)(i,_j)); }
    template<class E> matrix_column& operator=(const vector_expression<E>& ve) {
        vector<value_type> tmp(ve); for(size_type i=0; i!=size(); ++i) { (*this)(i)=tmp(i); } return *this; }
    matrix_column& operator=(const matrix_column& o) {
        vector<value_type> tmp(o); for(size_type i=0; i!=size(); ++i) { (*this)(i)=tmp(i); } return *this; }
};
template<class T> matrix_row< matrix<T> > row(matrix<T>& m, std::size_t i) { return matrix_row< matrix<T> >(m,i); }
template<class T> matrix_row< const matrix<T> > row(const matrix<T>& m, std::size_t i) { return matrix_row< const matrix<T> >(m,i); }
template<class T> matrix_column< matrix<T> > column(matrix<T>& m, std::size_t j) { return matrix_column< matrix<T> >(m,j); }
template<class T> matrix_column< const matrix<T> > column(const matrix<T>& m, std::size_t j) { return matrix_column< const matrix<T> >(m,j); }

} } }
#endif
