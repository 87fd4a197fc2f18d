// TEST INFRASTRUCTURE — cv::line_descriptor::KeyLine with the field order of opencv_contrib
// (Thirdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp:107-146 in the reference tree)
#pragma once
#include "core/core.hpp"
namespace cv { namespace line_descriptor {
struct KeyLine {
  float angle;
  int class_id;
  int octave;
  Point2f pt;
  float response;
  float size;
  float startPointX, startPointY, endPointX, endPointY;
  float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
  float lineLength;
  int numOfPixels;
};
}}  // namespace cv::line_descriptor
