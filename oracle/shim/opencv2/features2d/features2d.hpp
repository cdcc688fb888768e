// OpenCV shim (oracle test infrastructure): everything lives in opencv2/core/core.hpp
#include "opencv2/core/core.hpp"
