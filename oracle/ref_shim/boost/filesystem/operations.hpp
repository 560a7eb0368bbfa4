// Stand-in: see ../filesystem.hpp.  Test infrastructure only.
#pragma once
#include "../filesystem.hpp"
