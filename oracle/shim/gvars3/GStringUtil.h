#pragma once
#include "gvars3.h"
