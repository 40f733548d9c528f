from boa_b200.cli import run

run()
