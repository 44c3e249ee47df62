// Oracle tooling only: eager stand-in for boost/numeric/ublas/vector.hpp. See storage.hpp.
#ifndef ORACLE_SHIM_UBLAS_VECTOR_HPP
#define ORACLE_SHIM_UBLAS_VECTOR_HPP
#include <boost/numeric/ublas/storage.hpp>
namespace boost { namespace numeric { namespace ublas {

template<class T> class vector : public vector_expression< vector<T> > {
    std::vector<T> _data;
  public:
    typedef T value_type;
    typedef std::size_t size_type;
    typedef T& reference;
    typedef const T& const_reference;
    typedef typename std::vector<T>::iterator iterator;
    typedef typename std::vector<T>::const_iterator const_iterator;
    vector() : _data() { }
    explicit vector(size_type n) : _data(n) { }
    vector(size_type n, const T& t) : _data(n,t) { }
    vector(const vector<T>& v) : _data(v._data) { }
    template<class E> vector(const vector_expression<E>& ve) : _data() {
        const E& e=ve(); _data.reserve(e.size()); for(size_type i=0; i!=e.size(); ++i) { _data.push_back(T(e(i))); } }
    vector<T>& operator=(const vector<T>& v) { _data=v._data; return *this; }
    template<class E> vector<T>& operator=(const vector_expression<E>& ve) {
        vector<T> tmp(ve); _data.swap(tmp._data); return *this; }
    size_type size() const { return _data.size(); }
    bool empty() const { return _data.empty(); }
    void resize(size_type n, bool preserve=true) { _data.resize(n); }
    void clear() { for(size_type i=0; i!=_data.size(); ++i) { _data[i]=T(); } }
    const T& operator[](size_type i) const { return _data[i]; }
    T& operator[](size_type i) { return _data[i]; }
    const T& operator()(size_type i) const { return _data[i]; }
    T& operator()(size_type i) { return _data[i]; }
    iterator begin() { return _data.begin(); }
    iterator end() { return _data.end(); }
    const_iterator begin() const { return _data.begin(); }
    const_iterator end() const { return _data.end(); }
    template<class E> vector<T>& operator+=(const vector_expression<E>& ve) {
        const E& e=ve(); for(size_type i=0; i!=size(); ++i) { _data[i]+=e(i); } return *this; }
    template<class E> vector<T>& operator-=(const vector_expression<E>& ve) {
        const E& e=ve(); for(size_type i=0; i!=size(); ++i) { _data[i]-=e(i); } return *this; }
    template<class S> typename std::enable_if<std::is_convertible<S,T>::value,vector<T>&>::type
    operator*=(const S& s) { for(size_type i=0; i!=size(); ++i) { _data[i]*=s; } return *this; }
    template<class S> typename std::enable_if<std::is_convertible<S,T>::value,vector<T>&>::type
    operator/=(const S& s) { for(size_type i=0; i!=size(); ++i) { _data[i]/=s; } return *this; }
    void swap(vector<T>& o) { _data.swap(o._data); }
};

template<class E1, class E2>
vector<typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type>
operator+(const vector_expression<E1>& a, const vector_expression<E2>& b) {
    typedef typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type R;
    const E1& x=a(); const E2& y=b(); vector<R> r(x.size());
    for(std::size_t i=0; i!=x.size(); ++i) { r[i]=x(i)+y(i); } return r; }
template<class E1, class E2>
vector<typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type>
operator-(const vector_expression<E1>& a, const vector_expression<E2>& b) {
    typedef typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type R;
    const E1& x=a(); const E2& y=b(); vector<R> r(x.size());
    for(std::size_t i=0; i!=x.size(); ++i) { r[i]=x(i)-y(i); } return r; }
template<class E> vector<typename E::value_type> operator-(const vector_expression<E>& a) {
    const E& x=a(); vector<typename E::value_type> r(x.size());
    for(std::size_t i=0; i!=x.size(); ++i) { r[i]=-x(i); } return r; }
template<class S, class E>
typename std::enable_if<std::is_convertible<S,typename E::value_type>::value,
    vector<typename promote_traits<S,typename E::value_type>::promote_type> >::type
operator*(const S& s, const vector_expression<E>& a) {
    typedef typename promote_traits<S,typename E::value_type>::promote_type R;
    const E& x=a(); vector<R> r(x.size());
    for(std::size_t i=0; i!=x.size(); ++i) { r[i]=s*x(i); } return r; }
template<class E, class S>
typename std::enable_if<std::is_convertible<S,typename E::value_type>::value,
    vector<typename promote_traits<typename E::value_type,S>::promote_type> >::type
operator*(const vector_expression<E>& a, const S& s) {
    typedef typename promote_traits<typename E::value_type,S>::promote_type R;
    const E& x=a(); vector<R> r(x.size());
    for(std::size_t i=0; i!=x.size(); ++i) { r[i]=x(i)*s; } return r; }
template<class E, class S>
typename std::enable_if<std::is_convertible<S,typename E::value_type>::value,
    vector<typename promote_traits<typename E::value_type,S>::promote_type> >::type
operator/(const vector_expression<E>& a, const S& s) {
    typedef typename promote_traits<typename E::value_type,S>::promote_type R;
    const E& x=a(); vector<R> r(x.size());
    for(std::size_t i=0; i!=x.size(); ++i) { r[i]=x(i)/s; } return r; }

template<class E> typename E::value_type sum(const vector_expression<E>& a) {
    typedef typename E::value_type R; const E& x=a(); R t=R(0);
    for(std::size_t i=0; i!=x.size(); ++i) { t+=x(i); } return t; }

template<class E1, class E2>
typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type
inner_prod(const vector_expression<E1>& a, const vector_expression<E2>& b) {
    typedef typename promote_traits<typename E1::value_type,typename E2::value_type>::promote_type R;
    const E1& x=a(); const E2& y=b(); R t=R(0);
    for(std::size_t i=0; i!=x.size(); ++i) { t+=x(i)*y(i); } return t; }

} } }
#endif
