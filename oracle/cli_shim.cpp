// cli_shim.cpp -- TEST INFRASTRUCTURE.  Builds the reference's own CLI (main.cpp,
// compiled where it lies under /root/reference, unmodified) against the B200 engine's
// `Recommender` class: include/sr_recommender.hpp is seen first and claims the include
// guard of the reference's Recommender.h, so main.cpp's `#include "Recommender.h"`
// becomes a no-op and every Recommender call in main.cpp:59-82 lands in
// libsr_recommender.so.  See INTEGRATION.md for the one-line change a maintainer makes.
#include "sr_recommender.hpp"

#include "main.cpp"
