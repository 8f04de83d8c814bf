// What cmake would generate from the reference's config.h.in. Test infrastructure only.
#pragma once
#define PROJECT "Stage"
#define VERSION "4.3.0"
#define APIVERSION "4.3"
#define INSTALL_PREFIX "/nonexistent"
#define PLUGIN_PATH "/nonexistent/lib"
