"""2-GPU NCCL tests of the data-parallel gradient mean: ``parallel.GradReducer`` over NVLink with fp32 and
bf16 buckets, and the per-layer hand-off from inside ``functions.EncoderFn.backward`` on the real model -- the averaged
gradients must equal the mean of the two ranks' single-process gradients.  Skipped on a box with fewer than 2 devices
(run with ``gpurun --gpus 2``; log in ``profiles/``)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import modcr_b200  # noqa: F401

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from modcr_b200 import parallel
    r, local, w = parallel.init_from_env("nccl")
    assert (r, w) == (rank, world)
    torch.cuda.set_device(local)
    return parallel, torch.device("cuda", local)


def _reducer_worker(rank, world, port, q):
    try:
        parallel, dev = _init(rank, world, port)
        g = torch.Generator().manual_seed(5)
        shapes = [(768, 768), (3072, 768), (768,), (30567, 768), (2,), (1536, 768), (7, 3)]
        base = [torch.randn(*s, generator=g) for s in shapes]
        out = {}
        for comm, tol in ((torch.bfloat16, 1e-2), (None, 1e-6)):
            ps = [torch.nn.Parameter(torch.zeros(*s, device=dev)) for s in shapes]
            red = parallel.GradReducer(ps, bucket_mb=8, comm_dtype=comm)
            # half of the tensors arrive through the hand-off (as from EncoderFn.backward), half through autograd hooks
            red.handoff([(p, (b.to(dev) * (rank + 1)).contiguous()) for p, b in zip(ps[:3], base[:3])])
            sum((p * (b.to(dev) * (rank + 1))).sum() for p, b in zip(ps[3:], base[3:])).backward()
            red.finish()
            torch.cuda.synchronize()
            mean_factor = sum(r + 1 for r in range(world)) / world
            worst = max(float((p.grad.cpu() - b * mean_factor).abs().max() / b.abs().max()) for p, b in zip(ps, base))
            out[str(comm)] = (worst, tol, red.launched)
            assert worst < tol, (comm, worst)
            red.remove()
        dist.barrier()
        q.put((rank, True, out))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc() + str(e)))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def _model_worker(rank, world, port, q):
    """Real model, both encoders trainable: DP step with the per-layer hand-off vs. the mean of the single-process
    gradients of both ranks' batches."""
    try:
        parallel, dev = _init(rank, world, port)
        import params
        import restate
        import statedict
        from modcr_b200 import synth
        from modcr_b200.modeling_bert import BertImgModel
        from modcr_b200.modeling_utils import BertConfig
        from modcr_b200.modeling_vcr_chunkalign_v10 import ChunkAlign_CLS_enc4_align, SeqBertImgModel
        cfg_o = restate.default_cfg()
        cfg = BertConfig(vocab_size=cfg_o["vocab_size"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        m = ChunkAlign_CLS_enc4_align(BertImgModel(cfg), SeqBertImgModel(cfg), 4)
        m.load_state_dict(params.fill_params(statedict.materialise(statedict.chunkalign_keys(cfg_o)), 2), strict=True)
        m.to(dev).train()
        prefix = synth.make_prefix(12, 4, 768, seed=9).to(dev).requires_grad_(True)
        plist = list(m.parameters()) + [prefix]

        def step(seed_rank):
            b = synth.make_batch(B=1, T=24, R=7, seed=3, rank=seed_rank)
            bd = synth.to_device(b, dev)
            N = b["input_ids"].shape[0]
            out = m(bd["input_ids"], bd["img_feat"], input_mask=bd["input_mask"], label=bd["label"],
                    token_type_ids=bd["token_type_ids"],
                    encoder_history_states=[prefix[i].unsqueeze(0).expand(N, -1, -1) for i in range(12)],
                    offsets=b["offsets"], chunk_attention_mask=bd["chunk_attention_mask"],
                    gather_index=b["gather_index"], align_pos=bd["align_pos"], total_label=bd["total_label"])
            (out[0] + out[2]).backward()

        def grab():
            gs = [None if p.grad is None else p.grad.detach().clone() for p in plist]
            for p in plist:
                p.grad = None
            return gs
        # single-process gradients of both ranks' batches (no reducer active) -> exact mean
        assert parallel.active_reducer() is None
        singles = []
        for r in range(world):
            step(r)
            singles.append(grab())
        want = [None if a is None else sum(s[i] for s in singles) / world for i, a in enumerate(singles[0])]
        res = {}
        # fp32: atomics order only; the last case issues the collectives at the attention-backward windows
        for comm, tol, windows in ((None, 1e-4, False), (torch.bfloat16, 1e-2, False), (torch.bfloat16, 1e-2, True)):
            red = parallel.GradReducer(plist, bucket_mb=25, comm_dtype=comm, windows=windows)
            assert parallel.active_reducer() is red
            step(rank)
            red.finish()
            torch.cuda.synchronize()
            got = grab()
            worst, n = 0.0, 0
            for gv, wv in zip(got, want):
                assert (gv is None) == (wv is None)
                if gv is None:
                    continue
                n += 1
                worst = max(worst, float((gv - wv).norm() / wv.norm().clamp_min(1e-20)))
            res[str(comm) + ("+windows" if windows else "")] = (worst, tol, red.launched, n)
            assert worst < tol, (comm, worst)
            assert red.launched >= 20          # per-layer buckets, not one per encoder
            # a second step re-uses the reducer and its flat buffers
            step(rank)
            red.finish()
            torch.cuda.synchronize()
            got2 = grab()
            for gv, g2 in zip(got, got2):
                if gv is not None:         # same batch, same weights: equal up to the order of the atomic reductions
                    assert float((gv - g2).norm() / gv.norm().clamp_min(1e-20)) < 1e-3
            red.remove()
        dist.barrier()
        q.put((rank, True, res))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc() + str(e)))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def _run(worker):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
    print(res)
    assert all(ok for _, ok, _ in res), res


def test_grad_reducer_nccl_fp32_and_bf16_buckets():
    _run(_reducer_worker)


def test_model_data_parallel_step_equals_mean_of_single_process_gradients():
    _run(_model_worker)
