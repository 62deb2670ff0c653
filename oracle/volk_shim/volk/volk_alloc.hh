// TEST INFRASTRUCTURE ONLY (oracle). The reference only needs volk::vector
// (include/SoftFM.h:31, include/MultipathFilter.h:28); alignment is irrelevant for the
// scalar shim.
#ifndef ORACLE_VOLK_ALLOC_SHIM_HH
#define ORACLE_VOLK_ALLOC_SHIM_HH
#include <vector>
#include <volk/volk.h>
namespace volk {
template <class T> using vector = std::vector<T>;
}
#endif
