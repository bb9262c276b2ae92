#ifndef INCLUDE_GUARD
#define INCLUDE_GUARD

#define PROJECT_VERSION "0.2.3"

#endif