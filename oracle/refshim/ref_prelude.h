// ORACLE SUPPORT (test infrastructure).  Force-included (-include) in front of every
// reference translation unit compiled into oracle/_ref/libcrt_ref.so.  It supplies what the
// reference gets implicitly from MSVC (oracle/refshim/README.md lists each item with the
// reference line that needs it); it changes no arithmetic.
#pragma once
#include <optional>
#include <queue>
#include <unordered_map>
#include <iomanip>
#include <sstream>
#include <cmath>
// MSVC's <cmath> exposes ::sqrtf/::powf in namespace std (Shapes.h:512, Sampling.h:201,209, helpers.h:224)
namespace std { using ::sqrtf; using ::powf; }
#include "pch.h"  // the reference's own pch.h (found through -I /root/reference); MSVC force-includes it
// helpers.h:194,204 spell a template parameter `float` (accepted by MSVC, ill-formed elsewhere) and
// helpers.h:218 calls an undeclared `Float(...)`.  Renaming the token while that one header is read turns
// the parameter into an ordinary identifier; outside the two templates the alias is plain float.
typedef float Float;
typedef float crt_ref_f32;
#define float crt_ref_f32
#include "ThirdParty/pbrv4/helpers.h"
#undef float
// Shapes.h:940 uses `vertex_available()` as a default argument inside the enclosing class, which gcc
// rejects for a nested class with default member initialisers (CWG 1397).  Same value, spelled out
// (Shapes.h:917-924: true, true, true, true, false).
#define vertex_available() vertex_available{true, true, true, true, false}
