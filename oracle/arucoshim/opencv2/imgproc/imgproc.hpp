// oracle/arucoshim: everything lives in core.hpp
#pragma once
#include "../core/core.hpp"
