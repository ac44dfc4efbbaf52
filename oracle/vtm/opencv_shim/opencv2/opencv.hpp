// ORACLE BUILD SHIM: see core.hpp
#pragma once
#include "core.hpp"
