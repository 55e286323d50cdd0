// Forwarding header: lets sources written against supranational/sppark's <ff/batch_inversion.hpp> compile against
// libsppark_b200.so (include/sppark_b200.hpp has the same-named templates over the C ABI;
// INTEGRATION.md section 4).  Include one of the <ff/*.hpp> field headers (or define FEATURE_*) first,
// as the reference's own translation units do.
#pragma once
#ifndef SPPARK_B200_NO_DROPIN_DECLS
# define SPPARK_B200_NO_DROPIN_DECLS
#endif
#include "../../sppark_b200.hpp"
