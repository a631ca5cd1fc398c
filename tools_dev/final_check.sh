# round-end style validation on one GPU: build check, smoke, GPU tests, default bench
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout -k 10 300 python bench.py 2>&1 | grep -E "^\{|Error|Traceback" | cut -c1-400
timeout -k 10 60 python bench.py --impl reference | cut -c1-200
