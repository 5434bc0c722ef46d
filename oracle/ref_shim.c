/* TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
 *
 * Link shim that lets the UNMODIFIED reference sources under
 * /root/reference/software build without the Intel AAL service
 * (HelloALINLB.cpp cannot be built here: aalsdk is absent).
 *
 * The reference's bwa.c / fastmap.c `extern` six globals that
 * HelloALINLB.cpp:65-72 defines (the HARP shared-workspace pointers).
 * The oracle build uses -DUSE_SW, so nothing ever waits on them, but the
 * symbols must exist and point at real storage (bwa.c:307 stores to
 * *handshake unconditionally).
 */
#include <stdlib.h>

static unsigned int ref_handshake_word, ref_read_size_word;
static unsigned long int ref_in_buf[1 << 17], ref_out_buf[1 << 17]; /* 1 MB each, HelloALINLB.cpp:59-63 */

unsigned int *SPL_BWT_ref = 0;
unsigned int *SPL_CNT_table = 0;
unsigned int *read_size = &ref_read_size_word;
unsigned int *handshake = &ref_handshake_word;
unsigned long int *SPL_BWT_input = ref_in_buf;
unsigned long int *SPL_BWT_output = ref_out_buf;

#ifdef REF_SHIM_MAIN
int top_main(int argc, char *argv[]); /* top.c:63 */
int main(int argc, char *argv[]) { return top_main(argc, argv); }
#endif
