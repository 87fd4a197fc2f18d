// TEST INFRASTRUCTURE — C entry point over the reference's own ORB_SLAM2::LineIterator
// (/root/reference/add_src/lineIterator.cpp, compiled unchanged; recipe: oracle/Makefile, output oracle/_ref/).
// Frame::AssignFeaturesToGridForLine (Frame.cc:286-309) walks every key line through the 64 x 48 grid with it.
#include <cstdint>
#include <utility>

#include "lineIterator.h"

extern "C" int ref_line_iterator(double x1, double y1, double x2, double y2, int32_t* xy, int cap) {
  ORB_SLAM2::LineIterator it(x1, y1, x2, y2);
  std::pair<int, int> p;
  int n = 0;
  while (it.getNext(p)) {
    if (n < cap) {
      xy[2 * n] = p.first;
      xy[2 * n + 1] = p.second;
    }
    ++n;
  }
  return n;
}
