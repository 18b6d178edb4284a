#pragma once
#include "image.h"
