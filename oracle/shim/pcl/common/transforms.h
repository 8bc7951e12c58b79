// ORACLE — TEST INFRASTRUCTURE ONLY. Stand-in header; see pcl/point_types.h in this shim.
#include "../point_types.h"
