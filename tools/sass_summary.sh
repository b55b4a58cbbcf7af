# regenerates profiles/<tag>_sass_summary.txt from the built library (no GPU needed); $1 = tag (default r02)
tag=${1:-r02}
so=agarcl_b200/libagarcl_b200.so
out=profiles/${tag}_sass_summary.txt
cuobjdump -sass $so 2>/dev/null > /tmp/agarcl_all.sass
{
echo "# SASS summary of $so (sm_100a), HEAD of round 2"
echo "# produced by: bash tools/sass_summary.sh  (cuobjdump -sass | grep -c <mnemonic>;  python tools/sass_size.py)"
echo
echo "## mnemonic counts over all kernels of the library (k_step, k_order, k_selftest_sort, k_obs<int32/int16>, k_ram, k_reset, k_flags, k_pack<..>)"
for m in 'UBLKCP.G.S' 'UBLKCP.S.G' 'SYNCS' 'REDUX' 'ATOMS' 'ATOMG' 'REDG\|RED\.' 'DFMA' 'MUFU' 'BAR.SYNC\|BAR.ARV' 'STL' 'LDL' 'UTC.*MMA\|HMMA\|IMMA\|QGMMA'; do
  printf "%-28s %s\n" "$m" "$(grep -c "$m" /tmp/agarcl_all.sass)"
done
echo
echo "UBLKCP.G.S = cp.async.bulk shared->global (observation zero stream, channel-0 rows, pellet write-back); UBLKCP.S.G = bulk global->shared (pellet array);"
echo "SYNCS = mbarrier (bulk-load completion); REDUX = warp reductions; no tensor-core mnemonic is expected: nothing on the path is a contraction."
echo "STL / LDL of k_step alone: $(awk '/Function : /{f=0} /Function : .*k_stepENS/{f=1} f' /tmp/agarcl_all.sass | grep -c 'STL') / $(awk '/Function : /{f=0} /Function : .*k_stepENS/{f=1} f' /tmp/agarcl_all.sass | grep -c 'LDL') (round 1: 809 together; the spilled context moved to shared memory, DESIGN 3)"
echo
echo "## k_step code size by source function (instruction-cache argument: L1.5 I-cache 32 KB, L0 ~6 KB per SM sub-partition)"
python tools/sass_size.py 2>/dev/null
} > $out
wc -l $out
