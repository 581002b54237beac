#!/bin/bash
# runscript_cuda.sh -- the CUDA leg of the reference's result sweeps.
#
# The reference's runscript_<method>.sh (runscript_mc_eur.sh, runscript_mc_amer.sh, runscript_mc_asia.sh,
# runscript_mc_eur_multi.sh, runscript_binom_embar.sh) write a header to results/results_<method>.csv, bake
# `comparison` from a tree run, and append one row per (N, flavour, threads/processes). This script appends the
# "CUDA" rows of the same sweeps -- same S0/r/sigma/T/E, same N lists, same header, Parallel = number of GPUs where the
# reference writes threads or processes -- so the CSVs keep one format. Run it from the repository root after
# `make all_bin` (or from a reference checkout whose bin/ holds these executables).
#
#   bash runscript_cuda.sh [binom_embar|mc_eur|mc_eur_multi|mc_amer|mc_asia|all]
#   PCF_GPUS="1 2 4 8"   GPU counts to sweep (default: 1)
#   PCF_NS="..."         override the N list (quick runs / tests)
#   RESULTS_DIR=results  where the CSVs live
set -e
RESULTS_DIR=${RESULTS_DIR:-results}
GPUS=(${PCF_GPUS:-1})
HEADER="Method,Payoff,S0,E,r,sigma,T,N,M,Parallel,Nr_of_assets,T_overall,T_calculation,Result,Abs_Error,Error"
S0=100; r=0.02; sigma=0.75; T=1          # runscript_mc_eur.sh:5-8 (identical in every run-script)
mkdir -p "${RESULTS_DIR}" include

start_csv() {  # keep an existing file (rows of the other flavours), create it with the reference's header otherwise
  [ -s "$1" ] || echo "${HEADER}" > "$1"
}
ns() { if [ -n "${PCF_NS}" ]; then echo ${PCF_NS}; else echo "$@"; fi; }
col14() { tr ',' '\t' | awk '{print $14}'; }

do_binom_embar() {
  local out=${RESULTS_DIR}/results_binom_embar.csv E=110; start_csv ${out}
  export PCF_COMPARISON=26.61224                     # runscript_binom_embar.sh:22
  for N in $(ns 100 800 1000 1600 3200 6400 8000 10000 32000 64000 80000 100000); do
    ./bin/binom_embar call ${S0} ${E} ${r} ${sigma} ${T} ${N} 1 >> ${out}
    ./bin/binom_vanilla_eur call ${S0} ${E} ${r} ${sigma} ${T} ${N} >> ${out}
    echo "CUDA binom N=${N} -- DONE"
  done
}
do_mc_eur() {
  local out=${RESULTS_DIR}/results_mc_eur.csv E=110; start_csv ${out}
  # runscript_mc_eur.sh:22-24: comparison = the European tree at N = 100000 (45 s on the CPU, milliseconds here)
  export PCF_COMPARISON=$(./bin/binom_vanilla_eur call ${S0} ${E} ${r} ${sigma} ${T} 100000 | col14)
  for N in $(ns 10000 80000 100000 160000 320000 640000 800000 1000000 1600000 3200000 6400000 8000000 10000000 16000000 32000000 64000000 80000000 100000000); do
    for g in ${GPUS[@]}; do
      ./bin/mc_eur call ${S0} ${E} ${r} ${sigma} ${T} ${N} ${g} >> ${out}
      echo "CUDA mc_eur N=${N}, gpus=${g} -- DONE"
    done
  done
}
do_mc_amer() {
  local out=${RESULTS_DIR}/results_mc_amer.csv E=110 M=200; start_csv ${out}
  # runscript_mc_amer.sh:23-25: comparison = the American tree at N = 10000
  export PCF_COMPARISON=$(./bin/binom_vanilla_amer call ${S0} ${E} ${r} ${sigma} ${T} 10000 | col14)
  for N in $(ns 10000 80000 100000 160000 320000 640000 800000 1000000 1600000 3200000 6400000 8000000 10000000); do
    for g in ${GPUS[@]}; do
      ./bin/mc_amer call ${S0} ${E} ${r} ${sigma} ${T} ${N} ${M} ${g} >> ${out}
      echo "CUDA mc_amer N=${N}, M=${M}, gpus=${g} -- DONE"
    done
  done
}
do_mc_asia() {
  local out=${RESULTS_DIR}/results_mc_asia.csv E=110 M=200; start_csv ${out}
  export PCF_COMPARISON=13.71                        # runscript_mc_asia.sh:75-79 (online calculator value)
  for N in $(ns 100000 160000 320000 640000 800000 1000000 1600000 3200000 6400000 8000000 10000000); do
    for g in ${GPUS[@]}; do
      ./bin/mc_asia call ${S0} ${E} ${r} ${sigma} ${T} ${N} ${M} ${g} >> ${out}
      echo "CUDA mc_asia N=${N}, M=${M}, gpus=${g} -- DONE"
    done
  done
}
do_mc_eur_multi() {
  local out=${RESULTS_DIR}/results_mc_eur_multi.csv; start_csv ${out}
  local assets=(4 10) compares=(11.90 11.60) rho=0.5   # runscript_mc_eur_multi.sh:5-16 ("from article")
  local S0=100 E=100 r=0.1 sigma=0.2 T=1
  for i in ${PCF_ASSET_IDX:-0}; do                      # runscript_mc_eur_multi.sh:73 runs index 0 only (`for i in 0 #1`)
    export PCF_COMPARISON=${compares[$i]}
    for N in $(ns 10000 80000 100000 160000 320000 640000 800000 1000000 1600000 3200000 6400000 8000000 10000000 16000000 32000000 64000000 80000000 100000000); do
      for g in ${GPUS[@]}; do
        ./bin/mc_eur_multi call ${S0} ${E} ${r} ${sigma} ${T} ${N} ${assets[$i]} ${rho} ${g} >> ${out}
        echo "CUDA mc_eur_multi N=${N}, assets=${assets[$i]}, gpus=${g} -- DONE"
      done
    done
  done
}

case "${1:-all}" in
  binom_embar) do_binom_embar ;;
  mc_eur) do_mc_eur ;;
  mc_amer) do_mc_amer ;;
  mc_asia) do_mc_asia ;;
  mc_eur_multi) do_mc_eur_multi ;;
  all) do_binom_embar; do_mc_eur; do_mc_eur_multi; do_mc_amer; do_mc_asia ;;   # order of runscript.sh:3-7
  *) echo "usage: $0 [binom_embar|mc_eur|mc_eur_multi|mc_amer|mc_asia|all]" >&2; exit 2 ;;
esac
