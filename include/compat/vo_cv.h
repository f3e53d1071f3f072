// vo_cv.h -- picks real OpenCV when its headers exist, otherwise the minimal shim.
#pragma once
#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>) && !defined(VO_FORCE_CVSHIM)
#include <opencv2/core.hpp>
#define VO_HAVE_OPENCV 1
#endif
#endif
#ifndef VO_HAVE_OPENCV
#include "cvshim.h"
#endif
