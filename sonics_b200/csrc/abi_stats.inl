// extern "C" entry points for the statistics / repeat-until-pass path.
extern "C" {
int sonics_evaluate(sonics_ctx *, const SonicsParams *, const void *, int, const int32_t *, int, int64_t, int64_t,
                    const double *, const uint16_t *, int, SonicsVerdict *, void *)
{
    return fail(SONICS_ERR_ARG, "sonics_evaluate: not built yet");
}
size_t sonics_genotype_work_bytes(int, int, int) { return 0; }
int sonics_genotype_batch(sonics_ctx *, const SonicsParams *, int, const int32_t *, const int32_t *, const int32_t *,
                          const float *, const uint32_t *, int, uint64_t, SonicsVerdict *, void *, size_t, void *)
{
    return fail(SONICS_ERR_ARG, "sonics_genotype_batch: not built yet");
}
}
