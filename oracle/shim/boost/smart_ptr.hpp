#ifndef ORACLE_SHIM_BOOST_SMART_PTR_HPP
#define ORACLE_SHIM_BOOST_SMART_PTR_HPP
#include <boost/shared_ptr.hpp>
#include <boost/scoped_ptr.hpp>
#endif
