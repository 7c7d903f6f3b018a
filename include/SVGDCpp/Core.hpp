/* Core.hpp — common includes of the facade (reference Core.hpp).  The CppAD typedefs and converters
 * of the reference (Core.hpp:29-79) have no counterpart: the accelerated path has no AD types. */
#ifndef SVGDCPP_B200_CORE_HPP
#define SVGDCPP_B200_CORE_HPP

#include <cmath>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../svgd_b200.h"
#include "Exceptions.hpp"
#include "MiniEigen.hpp"

/* Reference Core.hpp:83-106 prepares CppAD's thread allocator for OpenMP.  Kept as a no-op so
 * programs that call it before constructing a parallel SVGD still compile and run. */
inline void SetupForParallelMode() {}

template <typename A, typename B>
bool CompareVectorSizes(const A &a, const B &b)
{
    return a.size() == b.size();
}
#endif
