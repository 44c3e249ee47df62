// Oracle tooling only: stand-in for boost::iterator_facade / iterator_adaptor (subset used by the reference).
#ifndef ORACLE_SHIM_BOOST_ITERATOR_FACADE_HPP
#define ORACLE_SHIM_BOOST_ITERATOR_FACADE_HPP
#include <iterator>
#include <cstddef>
#include <type_traits>
namespace boost {
struct use_default { };
struct forward_traversal_tag : std::forward_iterator_tag { };
struct bidirectional_traversal_tag : std::bidirectional_iterator_tag { };
struct random_access_traversal_tag : std::random_access_iterator_tag { };

class iterator_core_access {
  public:
    template<class I> static typename I::reference dereference(const I& i) { return i.dereference(); }
    template<class I> static void increment(I& i) { i.increment(); }
    template<class I> static void decrement(I& i) { i.decrement(); }
    template<class I, class J> static bool equal(const I& i, const J& j) { return i.equal(j); }
    template<class I, class D> static void advance(I& i, D n) { i.advance(n); }
    template<class I, class J> static std::ptrdiff_t distance_from(const I& i, const J& j) { return -i.distance_to(j); }
    template<class I, class J> static std::ptrdiff_t distance_to(const I& i, const J& j) { return i.distance_to(j); }
};

namespace detail {
template<class Cat> struct std_category { typedef Cat type; };
template<> struct std_category<forward_traversal_tag> { typedef std::forward_iterator_tag type; };
template<> struct std_category<bidirectional_traversal_tag> { typedef std::bidirectional_iterator_tag type; };
template<> struct std_category<random_access_traversal_tag> { typedef std::random_access_iterator_tag type; };
template<class R> struct arrow_proxy { R r; R* operator->() { return &r; } };
template<class R> struct arrow_result { typedef arrow_proxy<R> type; static type make(R r) { type p={r}; return p; } };
template<class T> struct arrow_result<T&> { typedef T* type; static type make(T& r) { return &r; } };
}

template<class Derived, class Value, class Category, class Reference=Value&, class Difference=std::ptrdiff_t>
class iterator_facade {
  public:
    typedef typename std::remove_const<Value>::type value_type;
    typedef Reference reference;
    typedef typename detail::arrow_result<Reference>::type pointer;
    typedef Difference difference_type;
    typedef typename detail::std_category<Category>::type iterator_category;
  protected:
    Derived& derived() { return *static_cast<Derived*>(this); }
    const Derived& derived() const { return *static_cast<const Derived*>(this); }
  public:
    reference operator*() const { return iterator_core_access::dereference(derived()); }
    pointer operator->() const { return detail::arrow_result<Reference>::make(iterator_core_access::dereference(derived())); }
    reference operator[](difference_type n) const { Derived t(derived()); iterator_core_access::advance(t,n); return *t; }
    Derived& operator++() { iterator_core_access::increment(derived()); return derived(); }
    Derived operator++(int) { Derived t(derived()); iterator_core_access::increment(derived()); return t; }
    Derived& operator--() { iterator_core_access::decrement(derived()); return derived(); }
    Derived operator--(int) { Derived t(derived()); iterator_core_access::decrement(derived()); return t; }
    Derived& operator+=(difference_type n) { iterator_core_access::advance(derived(),n); return derived(); }
    Derived& operator-=(difference_type n) { iterator_core_access::advance(derived(),-n); return derived(); }
    Derived operator-(difference_type n) const { Derived t(derived()); iterator_core_access::advance(t,-n); return t; }
};
template<class D,class V,class C,class R,class F>
inline D operator+(const iterator_facade<D,V,C,R,F>& i, typename iterator_facade<D,V,C,R,F>::difference_type n) {
    D t(static_cast<const D&>(i)); t+=n; return t; }
template<class D,class V,class C,class R,class F>
inline D operator+(typename iterator_facade<D,V,C,R,F>::difference_type n, const iterator_facade<D,V,C,R,F>& i) {
    D t(static_cast<const D&>(i)); t+=n; return t; }

#define ORACLE_SHIM_FACADE_REL(op,expr) \
template<class D1,class V1,class C1,class R1,class F1,class D2,class V2,class C2,class R2,class F2> \
inline bool operator op(const iterator_facade<D1,V1,C1,R1,F1>& a, const iterator_facade<D2,V2,C2,R2,F2>& b) { \
    const D1& i=static_cast<const D1&>(a); const D2& j=static_cast<const D2&>(b); return expr; }
ORACLE_SHIM_FACADE_REL(==, iterator_core_access::equal(i,j))
ORACLE_SHIM_FACADE_REL(!=, !iterator_core_access::equal(i,j))
ORACLE_SHIM_FACADE_REL(<,  iterator_core_access::distance_to(i,j)>0)
ORACLE_SHIM_FACADE_REL(>,  iterator_core_access::distance_to(i,j)<0)
ORACLE_SHIM_FACADE_REL(<=, iterator_core_access::distance_to(i,j)>=0)
ORACLE_SHIM_FACADE_REL(>=, iterator_core_access::distance_to(i,j)<=0)
#undef ORACLE_SHIM_FACADE_REL
template<class D1,class V1,class C1,class R1,class F1,class D2,class V2,class C2,class R2,class F2>
inline std::ptrdiff_t operator-(const iterator_facade<D1,V1,C1,R1,F1>& a, const iterator_facade<D2,V2,C2,R2,F2>& b) {
    return iterator_core_access::distance_to(static_cast<const D2&>(b),static_cast<const D1&>(a)); }

namespace detail {
template<class T, class Default> struct dflt { typedef T type; };
template<class Default> struct dflt<use_default,Default> { typedef Default type; };
}

template<class Derived, class Base, class Value=use_default, class Category=use_default,
         class Reference=use_default, class Difference=use_default>
class iterator_adaptor
    : public iterator_facade<Derived,
        typename detail::dflt<Value,typename std::iterator_traits<Base>::value_type>::type,
        typename detail::dflt<Category,typename std::iterator_traits<Base>::iterator_category>::type,
        typename detail::dflt<Reference,
            typename std::conditional<std::is_same<Value,use_default>::value,
                typename std::iterator_traits<Base>::reference,
                typename detail::dflt<Value,int>::type&>::type>::type,
        typename detail::dflt<Difference,typename std::iterator_traits<Base>::difference_type>::type>
{
    friend class iterator_core_access;
  public:
    typedef Base base_type;
    iterator_adaptor() : _base() { }
    explicit iterator_adaptor(const Base& b) : _base(b) { }
    const Base& base() const { return _base; }
  protected:
    typedef iterator_adaptor<Derived,Base,Value,Category,Reference,Difference> iterator_adaptor_;
    const Base& base_reference() const { return _base; }
    Base& base_reference() { return _base; }
  private:
    typename iterator_adaptor::reference dereference() const { return *_base; }
    template<class O> bool equal(const O& o) const { return _base==o.base(); }
    void increment() { ++_base; }
    void decrement() { --_base; }
    void advance(typename iterator_adaptor::difference_type n) { _base+=n; }
    template<class O> typename iterator_adaptor::difference_type distance_to(const O& o) const { return o.base()-_base; }
    Base _base;
};
} // namespace boost
#endif
