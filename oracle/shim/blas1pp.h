// Empty stand-in: everything lives in lapackpp.h (oracle build shim).
#pragma once
#include "lapackpp.h"
