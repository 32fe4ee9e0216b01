/* oracle/det_pthread.c — TEST INFRASTRUCTURE ONLY.
 *
 * Link-time interposers (ld --wrap) that turn the reference's 8 real worker
 * threads (core_legacy/src/normal_distributions.c:221-253) into 8 inline calls
 * in worker-id order.  That is one legal schedule of the reference ("worker 0
 * finishes, then worker 1, ...") and makes every voxel see its points in
 * ascending point index — the canonical order SURVEY.md Appendix A7 defines,
 * since the threaded reference is itself run-to-run nondeterministic.
 */
#include <pthread.h>
#include <stddef.h>

int __wrap_pthread_create(pthread_t *thread, const pthread_attr_t *attr, void *(*start)(void *), void *arg) {
    (void)attr;
    if (thread) *thread = pthread_self();
    start(arg);
    return 0;
}

int __wrap_pthread_join(pthread_t thread, void **retval) {
    (void)thread;
    if (retval) *retval = NULL;
    return 0;
}
