// Stand-in for <boost/math/tr1.hpp> (Boost is absent in this image).
// The reference's util.hh only does `using namespace boost::math::tr1;` to get
// log1p & friends, which <cmath> already provides on this toolchain.
#pragma once
#include <cmath>
namespace boost { namespace math { namespace tr1 {} } }
