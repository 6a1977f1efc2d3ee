#!/bin/bash
# rebuild libb2of.so with ptxas -v and print the resource lines of the kernels matching $1
python -c "
from hackathonopticalflow_b200 import _lib
print(_lib.build(force=True, verbose=True))" 2>&1 | grep -A3 "Compiling.*$1" | grep -v "^--" 
