// Stand-in for libmolgrid (absent third-party dependency), see grid_maker.h in this directory.
#pragma once
#include "libmolgrid/grid_maker.h"
