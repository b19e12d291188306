/* oracle/shim/intrin.h -- maps MSVC <intrin.h> (Core/Math.h:5) to the GCC header. Oracle build only. */
#pragma once
#include <x86intrin.h>
