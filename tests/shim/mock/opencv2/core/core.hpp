// TEST INFRASTRUCTURE — mock of the few OpenCV types include/psl_orbslam_shim.hpp touches: the stand-in headers that also
// compile the reference's ORBextractor.cc (oracle/ref_shim), plus the line_descriptor::KeyLine layout.
#pragma once
#include "../../../../../oracle/ref_shim/opencv2/core/core.hpp"
