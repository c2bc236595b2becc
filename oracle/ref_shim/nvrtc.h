// TEST INFRASTRUCTURE: empty stand-in; the reference's utils_host.h includes <nvrtc.h> but the
// loader half does not use it.
#pragma once
