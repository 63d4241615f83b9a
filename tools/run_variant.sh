#!/bin/bash
# Runs a command with an experimental variant of the library swapped in:  tools/run_variant.sh <name> <cmd...>
name=$1; shift
cp veils_b200/libveils_cuda.so /tmp/prod.so; cp veils_b200/build/libveils_var_$name.so veils_b200/libveils_cuda.so
"$@"
cp /tmp/prod.so veils_b200/libveils_cuda.so
