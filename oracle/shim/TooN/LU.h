#pragma once
#include "TooN.h"
