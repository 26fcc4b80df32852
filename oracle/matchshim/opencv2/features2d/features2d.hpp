// oracle/matchshim forwarding header (test infrastructure): everything lives in opencv2/core/core.hpp
#pragma once
#include "../core/core.hpp"
