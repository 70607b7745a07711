#pragma once
#include "glm/glm.hpp"
