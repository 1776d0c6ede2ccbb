/* Compiled by tests/test_abi.py against the REFERENCE's own header (when /root/reference is present) and linked against
 * libstringzillas_b200.so: proves the drop-in at the source and link level. Running it needs a GPU; without one it only
 * exercises the host-side entry points and expects the scoring call to refuse loudly. */
#include <stdio.h>
#include <string.h>
#include <stringzillas/stringzillas.h>

int main(void) {
    char const *why = NULL;
    szs_device_scope_t scope = NULL;
    szs_levenshtein_distances_t engine = NULL;
    if (szs_device_scope_init_default(&scope, &why) != sz_success_k) return 2;
    if (szs_levenshtein_distances_init(0, 1, 1, 1, NULL, szs_capabilities(), &engine, &why) != sz_success_k) return 3;
    char const *blob = "kittensitting";
    sz_u32_t const q_offsets[2] = {0, 6}, c_offsets[2] = {6, 13};
    sz_sequence_u32tape_t queries = {blob, q_offsets, 1}, candidates = {blob, c_offsets, 1};
    sz_size_t distance = 999;
    sz_status_t status = szs_levenshtein_distances_u32tape(engine, scope, &queries, &candidates, &distance, 1, &why);
    printf("version %d.%d.%d caps 0x%x status %d distance %zu message %s\n", szs_version_major(), szs_version_minor(),
           szs_version_patch(), (unsigned)szs_capabilities(), (int)status, (size_t)distance, why ? why : "(none)");
    szs_levenshtein_distances_free(engine);
    szs_device_scope_free(scope);
    if (status == sz_success_k) return distance == 3 ? 0 : 4;
    return status == sz_missing_gpu_k ? 0 : 5; /* no GPU: must refuse, never fall back */
}
