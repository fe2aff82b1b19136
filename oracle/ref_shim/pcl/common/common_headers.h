#pragma once
#include "../../standin_misc.h"
