"""`python -m wgbs_tools_b200.cli <command> [args]` -- the wgbstools dispatcher (reference src/python/wgbs_tools.py:50-79)
for the four hot-path commands (+ init_genome, which builds the CpG dictionary they need)."""
import importlib
import os
import sys
import time

COMMANDS = ("bam2pat", "pat2beta", "homog", "segment", "init_genome", "view", "cview", "beta_to_blocks", "beta_to_table", "index")


def _noop(argv):
    """start-up only (interpreter, library, process group, one Context per rank, one barrier): what every command pays under torchrun before
    its own work -- tools/multigpu_cli.py subtracts it from the commands' wall times"""
    from . import dist as wd
    from .api import Context
    rank, world, local = wd.init_from_env()
    with Context(local) as ctx:
        ctx.sync()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def main():
    if len(sys.argv) >= 2 and sys.argv[1] == "noop":
        return _noop(sys.argv[2:])
    if len(sys.argv) < 2 or sys.argv[1] in ("-h", "--help") or sys.argv[1] not in COMMANDS:
        print("Usage: wgbstools_b200 <command> [<args>]\nCommands: " + ", ".join(COMMANDS), file=sys.stderr)
        sys.exit(1 if len(sys.argv) < 2 or sys.argv[1] not in ("-h", "--help") else 0)
    try:
        mod = importlib.import_module("wgbs_tools_b200." + {"cview": "view"}.get(sys.argv[1], sys.argv[1]))
        if os.environ.get("WGBS_TIMING"):
            # wall clock of the command's own work: the process group (under torchrun) exists before the clock starts
            from . import dist as wd
            rank, world, _ = wd.init_from_env()
            if world > 1:
                import torch.distributed as dist
                dist.barrier()                                      # NCCL builds its communicator at the first collective: not the command's work
            t0 = time.perf_counter()
            mod.main(sys.argv[2:])
            if world > 1:
                import torch.distributed as dist
                dist.barrier()
            if rank == 0:
                print(f"[wgbstools_b200] {sys.argv[1]}: {time.perf_counter() - t0:.3f} s after start-up ({world} rank(s))", file=sys.stderr)
            return
        mod.main(sys.argv[2:])
    except ValueError as e:                                         # IllegalArgumentError (utils_wgbs.py:47-51)
        print(f"Invalid input argument\n{e}", file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
