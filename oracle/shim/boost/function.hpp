// oracle/shim/boost/function.hpp -- TEST INFRASTRUCTURE ONLY: boost::function -> std::function.
#pragma once
#include <functional>
namespace boost { template <class S> using function = std::function<S>; }
