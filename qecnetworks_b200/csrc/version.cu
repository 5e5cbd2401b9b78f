// version.cu -- ABI version of libqec_b200.so
extern "C" int qec_b200_version(void) { return 100; }
