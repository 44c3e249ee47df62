// Oracle tooling only: eager stand-in for boost/numeric/ublas/matrix.hpp (row-major dense). See storage.hpp.
#ifndef ORACLE_SHIM_UBLAS_MATRIX_HPP
#define ORACLE_SHIM_UBLAS_MATRIX_HPP
#include <boost/numeric/ublas/vector.hpp>
namespace boost { namespace numeric { namespace ublas {

template<class T> class matrix : public matrix_expression< matrix<T> > {
    std::size_t _r, _c; std::vector<T> _data;
  public:
    typedef T value_type;
    typedef std::size_t size_type;
    matrix() : _r(0), _c(0), _data() { }
    matrix(size_type r, size_type c) : _r(r), _c(c), _data(r*c) { }
    matrix(size_type r, size_type c, const T& t) : _r(r), _c(c), _data(r*c,t) { }
    matrix(const matrix<T>& m) : _r(m._r), _c(m._c), _data(m._data) { }
    template<class E> matrix(const matrix_expression<E>& me) : _r(me().size1()), _c(me().size2()), _data() {
        const E& e=me(); _data.reserve(_r*_c);
        for(size_type i=0; i!=_r; ++i) { for(size_type j=0; j!=_c; ++j) { _data.push_back(T(e(i,j))); } } }
    matrix<T>& operator=(const matrix<T>& m) { _r=m._r; _c=m._c; _data=m._data; return *this; }
    template<class E> matrix<T>& operator=(const matrix_expression<E>& me) {
        matrix<T> tmp(me); _r=tmp._r; _c=tmp._c; _data.swap(tmp._data); return *this; }
    size_type size1() const { return _r; }
    size_type size2() const { return _c; }
    void resize(size_type r, size_type c, bool preserve=true) {
        matrix<T> tmp(r,c);
        if(preserve) { for(size_type i=0; i<r && i<_r; ++i) { for(size_type j=0; j<c && j<_c; ++j) { tmp(i,j)=(*this)(i,j); } } }
        _r=r; _c=c; _data.swap(tmp._data); }
    const T& operator()(size_type i, size_type j) const { return _data[i*_c+j]; }
    T& operator()(size_type i, size_type j) { return _data[i*_c+j]; }
    template<class E> matrix<T>& operator+=(const matrix_expression<E>& me) {
        const E& e=me(); for(size_type i=0; i!=_r; ++i) { for(size_type j=0; j!=_c; ++j) { (*this)(i,j)+=e(i,j); } } return *this; }
    template<class E> matrix<T>& operator-=(const matrix_expression<E>& me) {
        const E& e=me(); for(size_type i=0; i!=_r; ++i) { for(size_type j=0; j!=_c; ++j) { (*this)(i,j)-=e(i,j); } } return *this; }
    template<class S> typename std::enable_if<std::is_convertible<S,T>::value,matrix<T>&>::type
    operator*=(const S& s) { for(size_type k=0; k!=_data.size(); ++k) { _data[k]*=s; } return *this; }
    template<class S> typename std::enable_if<std::is_convertible<S,T>::value,matrix<T>&>::type
    operator/=(const S& s) { for(size_type k=0; k!=_data.size(); ++k) { _data[k]/=s; } return *this; }
    void clear() { for(size_type k=0; k!=_data.size(); ++k) { _data[k]=T(); } }
};

template<class E1, class E2>
matrix<typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type>
operator+(const matrix_expression<E1>& a, const matrix_expression<E2>& b) {
    typedef typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type R;
    const E1& x=a(); const E2& y=b(); matrix<R> r(x.size1(),x.size2());
    for(std::size_t i=0; i!=x.size1(); ++i) { for(std::size_t j=0; j!=x.size2(); ++j) { r(i,j)=x(i,j)+y(i,j); } } return r; }
template<class E1, class E2>
matrix<typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type>
operator-(const matrix_expression<E1>& a, const matrix_expression<E2>& b) {
    typedef typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type R;
    const E1& x=a(); const E2& y=b(); matrix<R> r(x.size1(),x.size2());
    for(std::size_t i=0; i!=x.size1(); ++i) { for(std::size_t j=0; j!=x.size2(); ++j) { r(i,j)=x(i,j)-y(i,j); } } return r; }
template<class E> matrix<typename E::value_type> operator-(const matrix_expression<E>& a) {
    const E& x=a(); matrix<typename E::value_type> r(x.size1(),x.size2());
    for(std::size_t i=0; i!=x.size1(); ++i) { for(std::size_t j=0; j!=x.size2(); ++j) { r(i,j)=-x(i,j); } } return r; }
template<class S, class E>
typename std::enable_if<std::is_convertible<S,typename E::value_type>::value,
    matrix<typename promote_traits<S,typename E::value_type>::promote_type> >::type
operator*(const S& s, const matrix_expression<E>& a) {
    typedef typename promote_traits<S,typename E::value_type>::promote_type R;
    const E& x=a(); matrix<R> r(x.size1(),x.size2());
    for(std::size_t i=0; i!=x.size1(); ++i) { for(std::size_t j=0; j!=x.size2(); ++j) { r(i,j)=s*x(i,j); } } return r; }
template<class E, class S>
typename std::enable_if<std::is_convertible<S,typename E::value_type>::value,
    matrix<typename promote_traits<typename E::value_type,S>::promote_type> >::type
operator*(const matrix_expression<E>& a, const S& s) {
    typedef typename promote_traits<typename E::value_type,S>::promote_type R;
    const E& x=a(); matrix<R> r(x.size1(),x.size2());
    for(std::size_t i=0; i!=x.size1(); ++i) { for(std::size_t j=0; j!=x.size2(); ++j) { r(i,j)=x(i,j)*s; } } return r; }
template<class E, class S>
typename std::enable_if<std::is_convertible<S,typename E::value_type>::value,
    matrix<typename promote_traits<typename E::value_type,S>::promote_type> >::type
operator/(const matrix_expression<E>& a, const S& s) {
    typedef typename promote_traits<typename E::value_type,S>::promote_type R;
    const E& x=a(); matrix<R> r(x.size1(),x.size2());
    for(std::size_t i=0; i!=x.size1(); ++i) { for(std::size_t j=0; j!=x.size2(); ++j) { r(i,j)=x(i,j)/s; } } return r; }

// matrix * vector
template<class E1, class E2>
vector<typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type>
prod(const matrix_expression<E1>& a, const vector_expression<E2>& b) {
    typedef typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type R;
    const E1& A=a(); const E2& x=b(); vector<R> r(A.size1());
    for(std::size_t i=0; i!=A.size1(); ++i) { R t=R(0); for(std::size_t k=0; k!=A.size2(); ++k) { t+=A(i,k)*x(k); } r[i]=t; }
    return r; }
// vector * matrix
template<class E1, class E2>
vector<typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type>
prod(const vector_expression<E1>& a, const matrix_expression<E2>& b) {
    typedef typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type R;
    const E1& x=a(); const E2& A=b(); vector<R> r(A.size2());
    for(std::size_t j=0; j!=A.size2(); ++j) { R t=R(0); for(std::size_t k=0; k!=A.size1(); ++k) { t+=x(k)*A(k,j); } r[j]=t; }
    return r; }
// matrix * matrix
template<class E1, class E2>
matrix<typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type>
prod(const matrix_expression<E1>& a, const matrix_expression<E2>& b) {
    typedef typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type R;
    const E1& A=a(); const E2& B=b(); matrix<R> r(A.size1(),B.size2());
    for(std::size_t i=0; i!=A.size1(); ++i) { for(std::size_t j=0; j!=B.size2(); ++j) {
        R t=R(0); for(std::size_t k=0; k!=A.size2(); ++k) { t+=A(i,k)*B(k,j); } r(i,j)=t; } }
    return r; }
template<class E> matrix<typename E::value_type> trans(const matrix_expression<E>& a) {
    const E& A=a(); matrix<typename E::value_type> r(A.size2(),A.size1());
    for(std::size_t i=0; i!=A.size1(); ++i) { for(std::size_t j=0; j!=A.size2(); ++j) { r(j,i)=A(i,j); } } return r; }

} } }
#endif
