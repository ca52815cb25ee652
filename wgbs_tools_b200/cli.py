"""`python -m wgbs_tools_b200.cli <command> [args]` -- the wgbstools dispatcher (reference src/python/wgbs_tools.py:50-79)
for the four hot-path commands (+ init_genome, which builds the CpG dictionary they need)."""
import importlib
import sys

COMMANDS = ("bam2pat", "pat2beta", "homog", "segment", "init_genome", "view", "cview", "beta_to_blocks", "beta_to_table", "index")


def main():
    if len(sys.argv) < 2 or sys.argv[1] in ("-h", "--help") or sys.argv[1] not in COMMANDS:
        print("Usage: wgbstools_b200 <command> [<args>]\nCommands: " + ", ".join(COMMANDS), file=sys.stderr)
        sys.exit(1 if len(sys.argv) < 2 or sys.argv[1] not in ("-h", "--help") else 0)
    try:
        importlib.import_module("wgbs_tools_b200." + {"cview": "view"}.get(sys.argv[1], sys.argv[1])).main(sys.argv[2:])
    except ValueError as e:                                         # IllegalArgumentError (utils_wgbs.py:47-51)
        print(f"Invalid input argument\n{e}", file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
