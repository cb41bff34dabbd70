// TEST INFRASTRUCTURE ONLY -- see <opencv2/opencv.hpp> of this directory tree
#pragma once
#include <opencv2/opencv.hpp>
