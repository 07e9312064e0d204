// host side of the banded-alignment operator (bkid_align.cuh)
#pragma once

int bkid_op_banded_align(bkid_ctx *c, int64_t n, const uint8_t *q, const uint32_t *q_off, const uint8_t *r, const uint32_t *r_off, int32_t w, int32_t *out)
{
  if (!c || n < 0 || w < 0 || w > 15 || (n > 0 && (!q_off || !r_off || !out))) return c ? fail(c, BKID_ERR_ARG, "bad banded-align arguments (0 <= w <= 15)") : BKID_ERR_ARG;
  cudaSetDevice(c->device);
  c->err.clear();
  if (n == 0) return 0;
  cudaStream_t st = c->st;
  size_t qb = q_off[n], rb = r_off[n];
  DBuf dq, dqo, dr, dro, dout;
  int rc = 0;
  if ((rc = dq.ensure(qb + 16, 0, st)) || (rc = dr.ensure(rb + 16, 0, st)) || (rc = dqo.ensure((size_t)(n + 1) * 4, 0, st)) || (rc = dro.ensure((size_t)(n + 1) * 4, 0, st)) ||
      (rc = dout.ensure((size_t)n * 4, 0, st))) {
    for (DBuf *b : {&dq, &dqo, &dr, &dro, &dout}) b->release();
    return fail(c, rc, g_last_cuda_err);
  }
  if (qb) cudaMemcpyAsync(dq.p, q, qb, cudaMemcpyHostToDevice, st);
  if (rb) cudaMemcpyAsync(dr.p, r, rb, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dqo.p, q_off, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dro.p, r_off, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st);
  BK_LAUNCH(op_banded_align, (unsigned)std::min<long long>((n + AL_WARPS - 1) / AL_WARPS, 148 * 8), AL_WARPS * 32, 0, st, dq.as<uint8_t>(), dqo.as<uint32_t>(), dr.as<uint8_t>(),
            dro.as<uint32_t>(), (long long)n, (int)w, dout.as<int32_t>());
  rc = sync_check(c);
  if (!rc) cudaMemcpy(out, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
  for (DBuf *b : {&dq, &dqo, &dr, &dro, &dout}) b->release();
  return rc;
}
