// ORACLE — TEST INFRASTRUCTURE ONLY. Empty stand-in: include/utils/common.h includes this
// header, but nothing on the path (src/graph/trg.cpp) uses a symbol from it.
