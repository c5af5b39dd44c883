/* Debug driver: runs the tcgen05 building-block self-test of liboetr_b200.so (dlopen) and prints the errors.
 * Build: gcc -O1 -o tools/selftest_main tools/selftest_main.c -ldl ; run under compute-sanitizer on the GPU box. */
#include <dlfcn.h>
#include <stdio.h>
int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "imagematching-oetr_b200/liboetr_b200.so";
    void* h = dlopen(path, RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    int (*st)(float*, int) = (int (*)(float*, int))dlsym(h, "oetr_selftest_tcgen05");
    const char* (*le)(void) = (const char* (*)(void))dlsym(h, "oetr_last_error");
    float errs[16];
    int rc = st(errs, 16);
    printf("rc=%d (%s)\n", rc, rc ? le() : "ok");
    for (int i = 0; i < 8; ++i) printf("errs[%d]=%g\n", i, errs[i]);
    return rc != 0;
}
