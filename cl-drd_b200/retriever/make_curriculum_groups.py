"""Top-200 run -> curriculum group file: the producer the reference lacks for
`NwayDataset.create_from_relT_most_semi_hard_file` (dataset/nway_dataset.py:213-261).

    python make_curriculum_groups.py --run_path top200.run --output_path groups.train.json --label_mode 9 \
        [--teacher_run_path reranked.run] [--qrels_path qrels.train.tsv] [--most_window 10,50] [--semi_window 50,200] [--seed 0]

--teacher_run_path: the file the reference's re-ranker writes for the same candidates (evaluation/reranking_evaluator.py
-> evaluation/utils.py:write_rankdata): the student's top-200 are put into the teacher's order before they are cut.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cldrd import curriculum as CU  # noqa: E402


def _pair(s):
    a, b = s.split(",")
    return int(a), int(b)


def get_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--run_path", required=True, help="qid\\tpid\\trank\\tscore, ranked (retrieve_top_passages.py --top_k 200)")
    p.add_argument("--output_path", required=True)
    p.add_argument("--label_mode", default="9", choices=sorted(CU.LABEL_MODE_SHAPES, key=int))
    p.add_argument("--teacher_run_path", default=None, help="qid\\tpid\\trank\\tscore sorted by the teacher's score (write_rankdata)")
    p.add_argument("--qrels_path", default=None, help="qid\\t0\\tpid\\trel (evaluation/retrieval_evaluator.py reads the same file)")
    p.add_argument("--most_window", default=None, type=_pair)
    p.add_argument("--semi_window", default=None, type=_pair)
    p.add_argument("--seed", default=0, type=int)
    p.add_argument("--skip_short", action="store_true", help="skip queries whose run cannot fill the groups instead of failing")
    return p.parse_args(argv)


def read_qrels(path):
    out = {}
    with open(path) as f:
        for line in f:
            a = line.strip().split("\t")
            if len(a) == 4 and int(a[3]) >= 1:
                out.setdefault(int(a[0]), []).append(int(a[2]))
    return out


def main(args):
    qids, lists = CU.read_run(args.run_path)
    if args.teacher_run_path:
        t_qids, t_lists = CU.read_run(args.teacher_run_path)
        lists = CU.rerank_with_teacher(qids, lists, t_qids, t_lists)
    qrels = read_qrels(args.qrels_path) if args.qrels_path else None
    ex = CU.groups_for_label_mode(qids, lists, args.label_mode, most_window=args.most_window,
                                  semi_window=args.semi_window, qrels=qrels, seed=args.seed,
                                  strict=not args.skip_short)
    n = CU.write_groups(args.output_path, ex)
    print(f"# queries in run = {len(qids)}, examples written = {n}")


if __name__ == "__main__":
    main(get_args())
