# usage: bash tools/sweep_lib.sh <tag> <lib.so> "VAR=val ..." ...   -- sweep_env.sh on an alternative build of the library
cd $GRAFT_REPO_ROOT; tag=$1; lib=$2; shift; shift
BP_LIB_PATH=$PWD/$lib bash tools/sweep_env.sh $tag "$@"
