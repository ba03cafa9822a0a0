// Stand-in for <boost/cstdint.hpp> (Boost is absent in this image).
// Oracle build support only: lets the reference's mtw.hh / ziggurat.hh compile.
#pragma once
#include <stdint.h>
namespace boost {
using ::int8_t; using ::int16_t; using ::int32_t; using ::int64_t;
using ::uint8_t; using ::uint16_t; using ::uint32_t; using ::uint64_t;
}
