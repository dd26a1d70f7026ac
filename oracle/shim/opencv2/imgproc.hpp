#include <opencv2/mini_cv.hpp>
