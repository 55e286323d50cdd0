// Forwarding header: lets sources written against supranational/sppark's <ff/bls12-377-fp2.hpp> compile against
// libsppark_b200.so (include/sppark_b200.hpp has the types and templates).
// Such sources define the extern "C" entry points themselves, so the C header's own declarations
// of those names are hidden.
#pragma once
#ifndef SPPARK_B200_NO_DROPIN_DECLS
# define SPPARK_B200_NO_DROPIN_DECLS
#endif
#include "../../sppark_b200.hpp"
