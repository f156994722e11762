"""Predictor base + SimplePredictorManager -- easynlp/core/predictor.py:65-79,181-229: run = postprocess(predict(preprocess)),
TSV in / TSV out by `input_schema` / `output_schema` / `append_cols`.  New: an output file ending in `.npy` is a
binary sink -- the (single) output column must hold float32 vectors (CLIPPredictor(feature_format="numpy")) and the rows are
written as one [rows, E] array; `append_cols` then go to `<output>.tsv` next to it, row-aligned."""
import math


class Predictor(object):
    def __init__(self, *args, **kwargs):
        pass

    def preprocess(self, in_data):
        raise NotImplementedError

    def predict(self, in_data):
        raise NotImplementedError

    def postprocess(self, result):
        raise NotImplementedError

    def run(self, in_data):
        return self.postprocess(self.predict(self.preprocess(in_data)))


class SimplePredictorManager(object):
    def __init__(self, predictor, input_file, input_schema, output_file, output_schema, append_cols, skip_first_line=False, batch_size=32):
        self.predictor = predictor
        self.input_schema = input_schema
        self.output_schema = output_schema
        self.append_cols = append_cols
        self.batch_size = batch_size
        self.output_file = output_file
        with open(input_file, "r", encoding="utf-8") as f:
            if skip_first_line:
                f.readline()
            self.data_lines = f.readlines()

    def _batches(self, cols):
        for i in range(math.ceil(len(self.data_lines) / self.batch_size)):
            rows = [dict(zip(cols, ln.rstrip("\n").split("\t"))) for ln in self.data_lines[i * self.batch_size:(i + 1) * self.batch_size]]
            yield rows, self.predictor.run(rows)

    def _run_npy(self, cols):
        import numpy as np
        out_cols = self.output_schema.split(",")
        if len(out_cols) != 1:
            raise ValueError("a .npy sink takes exactly one output column")
        vecs, side = [], []
        for rows, outs in self._batches(cols):
            for row, od in zip(rows, outs):
                v = od[out_cols[0]]
                if isinstance(v, str):
                    raise TypeError("the .npy sink needs array features: construct the predictor with feature_format='numpy'")
                vecs.append(np.asarray(v, dtype=np.float32))
                if self.append_cols:
                    side.append("\t".join(str(row[c]) for c in self.append_cols.split(",")))
        np.save(self.output_file, np.stack(vecs) if vecs else np.zeros((0, 0), np.float32))
        if self.append_cols:
            with open(self.output_file + ".tsv", "w", encoding="utf-8") as f:
                f.write("".join(ln + "\n" for ln in side))

    def run(self):
        cols = [c.split(":")[0] for c in self.input_schema.split(",")]
        if self.output_file.endswith(".npy"):
            return self._run_npy(cols)
        with open(self.output_file, "w", encoding="utf-8") as fout:
            for i in range(math.ceil(len(self.data_lines) / self.batch_size)):
                rows = [dict(zip(cols, ln.rstrip("\n").split("\t"))) for ln in self.data_lines[i * self.batch_size:(i + 1) * self.batch_size]]
                outs = self.predictor.run(rows)
                for row, od in zip(rows, outs):
                    rec = [str(od[c]) for c in self.output_schema.split(",")]
                    if self.append_cols:
                        rec += [str(row[c]) for c in self.append_cols.split(",")]
                    fout.write("\t".join(rec) + "\n")
