#pragma once
#include "renderer/Hitpoint.h"
