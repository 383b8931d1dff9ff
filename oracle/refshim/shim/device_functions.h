#pragma once
#include "shim_optix.h"
