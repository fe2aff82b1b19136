#pragma once
#include "../../standin_opencv.h"
