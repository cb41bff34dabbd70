// TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's Evaluator.h (Middlebury accuracy logging + highgui windows, out of scope:
// SURVEY.md section 2) so that FastGCStereo.h / PMStereoBase.h compile for oracle/_ref.  The optimiser classes only hold a pointer to an
// Evaluator and test it for null before every use (FastGCStereo.h:65,139,156,...); the oracle never sets one.
#pragma once
#include "TimeStamper.h"
#include <opencv2/opencv.hpp>
#include "StereoEnergy.h"
class Evaluator {
public:
    double lastAccuracy = 0;
    bool showProgress = false, saveProgress = false, printProgress = false;
    void evaluate(cv::Mat, cv::Mat, const StereoEnergy&, bool, bool, bool, int, int = 0) {}
    void start() {}
    void stop() {}
    std::string getSaveDirectory() { return "./"; }
};
