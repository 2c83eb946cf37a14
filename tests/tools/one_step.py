"""ONE alternating training step between cudaProfilerStart / Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tests/tools/one_step.py [av|m3ae]
Single stream, graphs off (every launch listed, no co-running kernels): set before the package is imported."""
import argparse
import contextlib
import io
import os
import sys

os.environ.setdefault("MLA_OVERLAP", "0")
os.environ.setdefault("MLA_OVERLAP_WGRAD", "0")
os.environ.setdefault("MLA_GRAPHS", "0")
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mla_b200  # noqa: E402
from mla_b200.main import SyntheticAVLoader, SyntheticTextImageLoader  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "av"
    dev = torch.device("cuda:0")
    if which == "av":
        args = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modulation="Normal", gs_flag=True, dynamic=True,
                                  lorb="base", modal3=False, clip=False)
        mla_b200.setup_seed(0)
        model = mla_b200.ModuleHolder(mla_b200.AVClassifier(args).apply(mla_b200.weight_init).to(dev))
        batches = [tuple(t.to(dev) for t in b) for b in SyntheticAVLoader(64, 2, seed=1).batches]
        gs = mla_b200.GSPlugin(force_projection=True)
    else:
        args = argparse.Namespace(dataset="Food101", fusion_method="concat", modulation="Normal", gs_flag=True, dynamic=True,
                                  lorb="m3ae", modal3=False, clip=False)
        mla_b200.setup_seed(0)
        model = mla_b200.ModuleHolder(mla_b200.M3AEClassifier(args).to(dev))
        batches = [tuple(t.to(dev) if torch.is_tensor(t) else t for t in b) for b in SyntheticTextImageLoader(64, 2, 1).batches]
        gs = mla_b200.GSPlugin()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)

    def run(n):
        with contextlib.redirect_stdout(io.StringIO()):
            return mla_b200.train_epoch(args, 0, model, dev, [batches[i % 2] for i in range(n)], opt, sch, gs_plugin=gs,
                                        gs_flag=True, av_alpha=0.55)
    run(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    e0.record()
    run(1)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("one %s step, single stream, no graphs: %.3f ms" % (which, e0.elapsed_time(e1)))


if __name__ == "__main__":
    main()
