// Oracle tooling only: stand-in for boost/logic/tribool.hpp (three-valued logic, Kleene semantics).
#ifndef ORACLE_SHIM_BOOST_TRIBOOL_HPP
#define ORACLE_SHIM_BOOST_TRIBOOL_HPP
namespace boost { namespace logic {
namespace detail { struct indeterminate_t { }; }
class tribool;
inline bool indeterminate(tribool x, detail::indeterminate_t dummy=detail::indeterminate_t());
typedef bool (*indeterminate_keyword_t)(tribool, detail::indeterminate_t);
class tribool {
  public:
    enum value_t { false_value, true_value, indeterminate_value } value;
    tribool() : value(false_value) { }
    tribool(bool b) : value(b?true_value:false_value) { }
    tribool(indeterminate_keyword_t) : value(indeterminate_value) { }
    // "safe bool" idiom as in Boost: implicitly convertible in boolean contexts, true only when definitely true
  private:
    struct dummy { void nonnull() { } };
    typedef void (dummy::*safe_bool)();
  public:
    operator safe_bool() const { return value==true_value ? &dummy::nonnull : 0; }
};
inline bool indeterminate(tribool x, detail::indeterminate_t) { return x.value==tribool::indeterminate_value; }
inline tribool operator!(tribool x) {
    return x.value==tribool::false_value ? tribool(true) : x.value==tribool::true_value ? tribool(false) : tribool(indeterminate); }
inline tribool operator&&(tribool x, tribool y) {
    if(x.value==tribool::false_value || y.value==tribool::false_value) return false;
    if(x.value==tribool::true_value && y.value==tribool::true_value) return true;
    return indeterminate; }
inline tribool operator&&(tribool x, bool y) { return y ? x : tribool(false); }
inline tribool operator&&(bool x, tribool y) { return x ? y : tribool(false); }
inline tribool operator&&(indeterminate_keyword_t, tribool x) { return !x ? tribool(false) : tribool(indeterminate); }
inline tribool operator&&(tribool x, indeterminate_keyword_t) { return !x ? tribool(false) : tribool(indeterminate); }
inline tribool operator||(tribool x, tribool y) {
    if(x.value==tribool::true_value || y.value==tribool::true_value) return true;
    if(x.value==tribool::false_value && y.value==tribool::false_value) return false;
    return indeterminate; }
inline tribool operator||(tribool x, bool y) { return y ? tribool(true) : x; }
inline tribool operator||(bool x, tribool y) { return x ? tribool(true) : y; }
inline tribool operator||(indeterminate_keyword_t, tribool x) { return x ? tribool(true) : tribool(indeterminate); }
inline tribool operator||(tribool x, indeterminate_keyword_t) { return x ? tribool(true) : tribool(indeterminate); }
inline tribool operator==(tribool x, tribool y) {
    if(indeterminate(x) || indeterminate(y)) return indeterminate;
    return (x && y) || (!x && !y); }
inline tribool operator==(tribool x, bool y) { return x==tribool(y); }
inline tribool operator==(bool x, tribool y) { return tribool(x)==y; }
inline tribool operator==(indeterminate_keyword_t, tribool x) { return tribool(indeterminate)==x; }
inline tribool operator==(tribool x, indeterminate_keyword_t) { return tribool(indeterminate)==x; }
inline tribool operator!=(tribool x, tribool y) {
    if(indeterminate(x) || indeterminate(y)) return indeterminate;
    return !((x && y) || (!x && !y)); }
inline tribool operator!=(tribool x, bool y) { return x!=tribool(y); }
inline tribool operator!=(bool x, tribool y) { return tribool(x)!=y; }
inline tribool operator!=(indeterminate_keyword_t, tribool x) { return tribool(indeterminate)!=x; }
inline tribool operator!=(tribool x, indeterminate_keyword_t) { return x!=tribool(indeterminate); }
} // namespace logic
using logic::tribool;
using logic::indeterminate;
} // namespace boost
#endif
