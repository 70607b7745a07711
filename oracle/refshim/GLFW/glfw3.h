// stub: only the GLFWwindow name is needed by Util/InputManager.h declarations
#pragma once
struct GLFWwindow;
